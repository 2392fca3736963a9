// Shared by the facade test programs: a shadow copy of the device image rebuilt from the copy regions the facade
// issued, and the per-frame dump (format "CADRF002") that tests/facade_dump.py parses and checks against the oracle.
#pragma once
#include <CadR/CadR.h>
#include "../../../include/cadr_b200.h"
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <vector>

namespace dump {

using namespace CadR;

// DataMemory base address -> bytes, kept up to date by DataStorage::uploadObserver
struct Shadow {
	std::map<uint64_t, std::vector<uint8_t>> seg;
	void sync(DataStorage& ds) { for(DataMemory* m : ds.dataMemoryList()) if(m->size() && !seg.count(m->deviceAddress())) seg[m->deviceAddress()].assign(m->size(), 0); }
	void apply(const cadr_copy_region* r, size_t n) {
		for(size_t i = 0; i < n; i++) {
			auto it = seg.upper_bound(r[i].dstAddr);
			if(it == seg.begin()) { fprintf(stderr, "region outside every DataMemory\n"); exit(3); }
			--it;
			if(r[i].dstAddr + r[i].bytes > it->first + it->second.size()) { fprintf(stderr, "region overruns its DataMemory\n"); exit(3); }
			memcpy(it->second.data() + (r[i].dstAddr - it->first), reinterpret_cast<const void*>(r[i].srcOffset), r[i].bytes);
		}
	}
	void attach(Renderer& r) { r.dataStorage().uploadObserver = [this, &r](const cadr_copy_region* regs, size_t n) { sync(r.dataStorage()); apply(regs, n); }; }
};

// what the application expects the processing pass to write for drawable i of a draw range:
// indirect = {indexCount, instanceCount, firstIndex, 0}, pointers = {vertices, indices, matrix list, drawable data}
using ExpectFn = std::function<void(Drawable& d, const DrawableGpuData& rec, uint32_t indirect[4], uint64_t pointers[4])>;

struct Writer {
	FILE* out = nullptr;
	template<typename T> void put(const T& v) { fwrite(&v, sizeof(T), 1, out); }
	void putBytes(const void* p, size_t n) { if(n) fwrite(p, 1, n, out); }

	// one frame: device image, flattened drawable list + culling records, frustum, draw ranges, expected Tier R
	// records, output regions and - with a device - what the GPU wrote (Tier R buffers, compacted Tier X buffers)
	void frame(Renderer& r, Shadow& shadow, size_t n, const Frustum& f, int frameIndex, const ExpectFn& expect)
	{
		shadow.sync(r.dataStorage());
		fwrite("CADRF002", 1, 8, out);
		put(uint32_t(frameIndex)); put(uint32_t(r.hasDevice()));
		put(uint32_t(shadow.seg.size()));
		for(auto& [base, bytes] : shadow.seg) { put(uint64_t(base)); put(uint64_t(bytes.size())); putBytes(bytes.data(), bytes.size()); }
		put(uint64_t(r.dataStorage().handleTableDeviceAddress())); put(uint32_t(r.dataStorage().handleLevel())); put(uint32_t(n));
		put(uint64_t(r.dataStorage().handleTable().highestHandle()));
		put(uint64_t(r.drawableBufferAddress()));
		putBytes(r.drawableStagingData(), n * 48);
		putBytes(r.cullStagingData(), n * 48);
		put(f);
		const auto& ranges = r.drawRanges();
		put(uint32_t(ranges.size()));
		for(size_t k = 0; k < ranges.size(); k++) {
			put(uint32_t(ranges[k].firstDrawable)); put(uint32_t(ranges[k].numDrawables));
			put(uint64_t(ranges[k].drawablePointersAddress - r.drawablePointersBufferAddress())); put(uint64_t(ranges[k].indirectOffset));
		}
		for(const DrawRange& dr : ranges)
			for(size_t i = 0; i < dr.numDrawables; i++) {
				uint32_t ind[4] = {0, 0, 0, 0};
				uint64_t ptr[4] = {0, 0, 0, 0};
				expect(dr.stateSet->getDrawable(i), dr.stateSet->drawableDataList()[i], ind, ptr);
				putBytes(ind, 16); putBytes(ptr, 32);
			}
		const CullResult& c = r.cullResult();
		put(uint32_t(c.numRanges));
		for(auto& reg : c.regions) putBytes(reg.data(), 16);
		if(r.hasDevice()) {
			std::vector<uint8_t> buf(n * 48);
			r.readDevice(buf.data(), r.drawIndirectBufferAddress(), n * 16); putBytes(buf.data(), n * 16);
			r.readDevice(buf.data(), r.drawablePointersBufferAddress(), n * 32); putBytes(buf.data(), n * 32);
			uint64_t cmdCap = 0, instCap = 0;
			for(auto& reg : c.regions) { cmdCap += reg[1]; instCap += reg[3]; }
			put(cmdCap); put(instCap);
			const size_t cb = cadr_b200_cull_counters_bytes(c.numRanges);
			std::vector<uint8_t> big(std::max<size_t>({size_t(cmdCap) * 32, size_t(instCap) * 4, cb, 16}));
			r.readDevice(big.data(), c.counters, cb); putBytes(big.data(), cb);
			if(cmdCap) {
				r.readDevice(big.data(), c.commands, cmdCap * 20); putBytes(big.data(), cmdCap * 20);
				r.readDevice(big.data(), c.pointers, cmdCap * 32); putBytes(big.data(), cmdCap * 32);
				r.readDevice(big.data(), c.tags, cmdCap * 8); putBytes(big.data(), cmdCap * 8);
			}
			if(instCap) { r.readDevice(big.data(), c.instances, instCap * 4); putBytes(big.data(), instCap * 4); }
		}
		fprintf(stderr, "frame %d: %zu drawables, %zu ranges, handle level %u (%llu handles), %zu DataMemory, staging in use %zu / pooled %zu, list upload %zu bytes\n",
		        frameIndex, n, ranges.size(), r.dataStorage().handleLevel(), (unsigned long long)r.dataStorage().handleTable().highestHandle(),
		        r.dataStorage().dataMemoryList().size(), r.stagingManager().numBlocksInUse(), r.stagingManager().numBlocksAvailable(),
		        r.lastDrawableUploadBytes());
	}
};

}  // namespace dump
