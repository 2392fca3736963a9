"""Child process of tests/test_external_gpu.py: a *consumer* that received an exported buffer as an inherited file
descriptor (what a Vulkan application would pass to VkImportMemoryFdInfoKHR), maps it through the C ABI and prints a
digest of its first `nbytes` bytes.  usage: external_consumer.py <fd> <allocated_bytes> <nbytes>"""
import hashlib
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cadr_b200  # noqa: E402

fd, allocated, nbytes = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
ctx = cadr_b200.Context(0)
addr = ctx.external_import_fd(fd, allocated)
out = np.empty(nbytes, dtype=np.uint8)
ctx.memcpy_d2h(out, addr)
ctx.sync()
print("digest", hashlib.sha256(out.tobytes()).hexdigest())
ctx.external_free(addr)
ctx.close()
