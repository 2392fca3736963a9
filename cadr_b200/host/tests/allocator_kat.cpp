// Known-answer test of CadR::RingSuballocator against placements chosen by the REFERENCE'S OWN allocator.
// Reads the command stream format of oracle/ref_alloc_probe.cpp from stdin and prints the same output format, so
// the pytest driver can diff the two byte for byte (tests/test_host_cpu.py).
#include <CadR/RingSuballocator.h>
#include <cstdio>
#include <map>

using namespace CadR;

int main()
{
	const uint64_t base = 0x10000000ull;
	unsigned long long bytes;
	if(scanf("%llu", &bytes) != 1) return 1;
	RingSuballocator<RingRecord> ring(base, bytes);
	std::map<long, RingRecord*> live;
	char c; long id; unsigned long long n;
	while(scanf(" %c %ld", &c, &id) == 2) {
		if(c == 'a') {
			if(scanf("%llu", &n) != 1) return 1;
			auto [addr, region] = ring.propose(n);
			if(region == 0) { printf("a %ld -1 0\n", id); continue; }
			live[id] = ring.commit(region, addr, n);
			printf("a %ld %lld %d\n", id, (long long)(addr - base), region);
		}
		else if(c == 'f') {
			auto it = live.find(id);
			if(it == live.end()) { printf("f %ld -1\n", id); continue; }   // the allocation had failed: nothing to free
			ring.release(it->second);
			live.erase(it);
			printf("f %ld %llu\n", id, (unsigned long long)ring.usedBytes());
		}
	}
	return 0;
}
