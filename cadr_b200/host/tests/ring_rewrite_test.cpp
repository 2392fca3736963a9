// Regression test for the upload runs of DataMemory under heavy per-frame rewriting: 40..60 allocations of 1 KiB in
// the 64 KiB first arena are re-allocated EVERY frame (realloc-on-write, DataAllocation.cpp:18-34), so more than 60 %
// of the arena is re-uploaded per frame, ring region 1 empties mid-frame (the allocator then swaps the roles of its two
// regions, CircularAllocationMemory.h:320-347) and the following wrap starts again at the beginning of the buffer.
// Every copy region recorded for the frame (DataMemory::recordUploads, DataMemory.cpp:400-446) must then
//   * have a sane size (<= the arena) and lie inside its arena,
//   * come from inside a staging block that the application's writes also stayed inside,
//   * carry exactly the bytes the application staged: a host mirror of the arenas, updated from the regions the way
//     the device would be, must hold every live allocation's last contents.
// Runs on an address-space-only renderer (CPU boxes) or a real device (argument: CUDA device index).
#include <CadR/CadR.h>
#include "../../../include/cadr_b200.h"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

using namespace CadR;

int main(int argc, char** argv)
{
	const int device = argc > 1 ? atoi(argv[1]) : Renderer::addressSpaceOnly;
	try {
		size_t frames = 0, regions = 0, maxRegionBytes = 0;
		for(size_t count = 38; count <= 62; count += 3) {
			for(size_t bytes : {size_t(1024), size_t(960), size_t(1500)}) {
				Renderer r(device);
				DataStorage& ds = r.dataStorage();
				std::map<uint64_t, std::vector<uint8_t>> mirror;          // arena base -> bytes as the device would hold them
				std::string problem;
				ds.uploadObserver = [&](const cadr_copy_region* regs, size_t n) {
					for(size_t i = 0; i < n; i++) {
						const cadr_copy_region& c = regs[i];
						regions++;
						if(c.bytes > maxRegionBytes) maxRegionBytes = size_t(c.bytes);
						const DataMemory* home = nullptr;
						for(const DataMemory* m : ds.dataMemoryList())
							if(c.dstAddr >= m->deviceAddress() && c.dstAddr + c.bytes <= m->deviceAddress() + m->size()) home = m;
						if(!home || c.bytes > home->size()) {
							problem = "copy region outside its arena: dst " + std::to_string(c.dstAddr) + " bytes " + std::to_string(c.bytes);
							return;
						}
						auto& img = mirror[home->deviceAddress()];
						img.resize(home->size());
						std::memcpy(img.data() + (c.dstAddr - home->deviceAddress()), reinterpret_cast<const void*>(c.srcOffset), c.bytes);
					}
				};
				std::vector<HandlelessAllocation> allocs;
				for(size_t i = 0; i < count; i++) allocs.emplace_back(ds);
				for(size_t frame = 0; frame < 40; frame++) {
					r.beginFrame();
					for(size_t i = 0; i < count; i++) {
						// a third of the frames rewrite everything, the others a moving 70 % window
						if(frame % 3 != 0 && (i + frame) % 10 >= 7) continue;
						StagingData sd = allocs[i].alloc(bytes);
						std::memset(sd.data(), int((frame * 131 + i * 7) & 0xff), bytes);
						*sd.data<uint64_t>() = (uint64_t(frame) << 32) | i;
					}
					r.executeCopyOperations();
					if(!problem.empty()) throw std::runtime_error(problem + " (frame " + std::to_string(frame) + ", " + std::to_string(count) + " x " + std::to_string(bytes) + ")");
					// every live allocation holds what was last written to it
					for(size_t i = 0; i < count; i++) {
						if(allocs[i].size() == 0) continue;
						const uint64_t addr = allocs[i].deviceAddress();
						const uint8_t* p = nullptr;
						for(auto& [base, img] : mirror)
							if(addr >= base && addr + bytes <= base + img.size()) p = img.data() + (addr - base);
						if(!p) throw std::runtime_error("allocation was never uploaded");
						const uint64_t tag = *reinterpret_cast<const uint64_t*>(p);
						const size_t f = size_t(tag >> 32), who = size_t(tag & 0xffffffffu);
						if(who != i || f > frame) throw std::runtime_error("allocation " + std::to_string(i) + " holds foreign data in frame " + std::to_string(frame));
						const uint8_t fill = uint8_t((f * 131 + i * 7) & 0xff);
						for(size_t b = 8; b < bytes; b++)
							if(p[b] != fill) throw std::runtime_error("allocation " + std::to_string(i) + " corrupted at byte " + std::to_string(b));
					}
					r.endFrame();
					frames++;
				}
				allocs.clear();
				r.executeCopyOperations();
				// usedBytes() is not checked: like the reference's counter (CircularAllocationMemory.h:574-639, pinned by
				// tests/golden/allocator_kat.json.gz) it loses the alignment padding in front of an allocation made right
				// after the region's newest allocation was released, so interleaved patterns may not return to 0
				for(const DataMemory* m : ds.dataMemoryList())
					if(!m->ringEmpty()) throw std::runtime_error("Not all memory was released (" + std::to_string(count) + " x " + std::to_string(bytes) + ")");
			}
		}
		// ---- writes after a transfer in the middle of a frame ------------------------------------------------------
		// The reference hands the SAME staging block out again for every write of a frame (DataAllocation.cpp:22-28), also
		// after executeCopyOperations() has transferred and released it: the second write then lands in recycled staging
		// memory and never reaches the device (an application that builds a big scene in batches inside one frame, or the
		// handle table's leaf that receives entries before and after the transfer).  The facade re-allocates instead.
		{
			Renderer r(device);
			DataStorage& ds = r.dataStorage();
			std::map<uint64_t, std::vector<uint8_t>> mirror;
			ds.uploadObserver = [&](const cadr_copy_region* regs, size_t n) {
				for(size_t i = 0; i < n; i++)
					for(const DataMemory* m : ds.dataMemoryList())
						if(regs[i].dstAddr >= m->deviceAddress() && regs[i].dstAddr + regs[i].bytes <= m->deviceAddress() + m->size()) {
							auto& img = mirror[m->deviceAddress()];
							img.resize(m->size());
							std::memcpy(img.data() + (regs[i].dstAddr - m->deviceAddress()), reinterpret_cast<const void*>(regs[i].srcOffset), regs[i].bytes);
						}
			};
			auto deviceByte = [&](uint64_t addr) -> int {
				for(auto& [base, img] : mirror) if(addr >= base && addr < base + img.size()) return img[addr - base];
				return -1;
			};
			r.beginFrame();
			std::vector<DataAllocation> objs;
			HandlelessAllocation a(ds);
			std::memset(a.alloc(256).data(), 0x11, 256);
			for(int i = 0; i < 100; i++) { objs.emplace_back(ds); std::memset(objs.back().alloc(64).data(), i, 64); }    // handle table entries 1..100
			r.executeCopyOperations();
			if(deviceByte(a.deviceAddress()) != 0x11) throw std::runtime_error("first write did not reach the device");
			StagingData again = a.alloc(256);                     // same frame, after the transfer
			if(!again.wasReallocated()) throw std::runtime_error("a block that was already transferred and released was handed out again");
			std::memset(again.data(), 0x22, 256);
			for(int i = 0; i < 100; i++) { objs.emplace_back(ds); std::memset(objs.back().alloc(64).data(), 100 + i, 64); }   // more entries into the same leaf table
			r.executeCopyOperations();
			r.endFrame();
			if(deviceByte(a.deviceAddress()) != 0x22) throw std::runtime_error("second write of the frame did not reach the device");
			// every handle of the (re-staged) leaf table resolves to its object's address
			const uint64_t root = ds.handleTableDeviceAddress();
			for(size_t i = 0; i < objs.size(); i++) {
				uint64_t entry = 0;
				for(int b = 7; b >= 0; b--) entry = (entry << 8) | uint64_t(deviceByte(root + 8 * objs[i].handle() + b) & 0xff);
				if(entry != objs[i].deviceAddress()) throw std::runtime_error("handle " + std::to_string(objs[i].handle()) + " does not resolve after a mid-frame transfer");
				if(deviceByte(objs[i].deviceAddress()) != int(i)) throw std::runtime_error("object data lost");
			}
		}
		// ---- a scene that keeps re-writing big allocations: old DataMemory objects empty out and are taken into service again
		// (DataStorage::setReuseEmptyDataMemories, on by default); with the reference's policy the same run only ever grows
		size_t memoriesWithReuse = 0, memoriesReference = 0;
		for(int reuse = 1; reuse >= 0; reuse--) {
			Renderer r(device);
			DataStorage& ds = r.dataStorage();
			ds.setReuseEmptyDataMemories(reuse != 0);
			std::map<uint64_t, std::vector<uint8_t>> mirror;
			ds.uploadObserver = [&](const cadr_copy_region* regs, size_t n) {
				for(size_t i = 0; i < n; i++)
					for(const DataMemory* m : ds.dataMemoryList())
						if(regs[i].dstAddr >= m->deviceAddress() && regs[i].dstAddr + regs[i].bytes <= m->deviceAddress() + m->size()) {
							auto& img = mirror[m->deviceAddress()];
							img.resize(m->size());
							std::memcpy(img.data() + (regs[i].dstAddr - m->deviceAddress()), reinterpret_cast<const void*>(regs[i].srcOffset), regs[i].bytes);
						}
			};
			const size_t count = 48, bytes = 3u << 20;               // 144 MiB live in 32 MiB memories, a quarter re-written per frame
			std::vector<HandlelessAllocation> allocs;
			for(size_t i = 0; i < count; i++) allocs.emplace_back(ds);
			for(size_t frame = 0; frame < 24; frame++) {
				r.beginFrame();
				for(size_t i = 0; i < count; i++) {
					if(frame != 0 && (i + frame) % 4 != 0) continue;
					StagingData sd = allocs[i].alloc(bytes);
					uint64_t* w = sd.data<uint64_t>();
					w[0] = (uint64_t(frame) << 32) | i; w[bytes / 8 - 1] = ~w[0];
				}
				r.executeCopyOperations();
				for(size_t i = 0; i < count; i++) {
					const uint64_t addr = allocs[i].deviceAddress();
					const uint8_t* p = nullptr;
					for(auto& [base, img] : mirror) if(addr >= base && addr + bytes <= base + img.size()) p = img.data() + (addr - base);
					if(!p) throw std::runtime_error("big allocation was never uploaded");
					const uint64_t head = *reinterpret_cast<const uint64_t*>(p), tail = *reinterpret_cast<const uint64_t*>(p + bytes - 8);
					if((head & 0xffffffffu) != i || tail != ~head) throw std::runtime_error("big allocation " + std::to_string(i) + " holds foreign data in frame " + std::to_string(frame));
				}
				r.endFrame();
			}
			(reuse ? memoriesWithReuse : memoriesReference) = ds.dataMemoryList().size();
		}
		if(memoriesWithReuse >= memoriesReference || memoriesWithReuse > 16)
			throw std::runtime_error("empty DataMemory objects are not taken into service again: " + std::to_string(memoriesWithReuse) + " vs " + std::to_string(memoriesReference));
		printf("ring_rewrite_test ok: %zu frames, %zu copy regions, largest %zu bytes; churn of 3 MiB allocations: %zu DataMemory objects with reuse, %zu with the reference's policy\n",
		       frames, regions, maxRegionBytes, memoriesWithReuse, memoriesReference);
		return 0;
	}
	catch(std::exception& e) {
		printf("ring_rewrite_test FAILED: %s\n", e.what());
		return 1;
	}
}
