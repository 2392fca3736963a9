// The scenarios of the reference's tests/DataAllocationTest.cpp (:55-326) against the facade: placement of
// HandlelessAllocations inside the first 64 KiB DataMemory and "memory returns to fully empty" after frees in
// eight different orders, including the region-1 / region-2 wrap.  Runs on an address-space-only renderer (CPU
// boxes) or on a real device (argument: CUDA device index).
#include <CadR/CadR.h>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <string>
#include <vector>

using namespace CadR;

static void verifyDataStorageEmpty(const DataStorage& ds)
{
	for(const DataMemory* m : ds.dataMemoryList())
		if(m->usedBytes() != 0 || !m->ringEmpty())
			throw std::runtime_error("Not all memory was released.");
}

static size_t slot(size_t s)
{
	// DataAllocationTest.cpp:104
	return (s == 0) ? 0 : (s <= 16) ? 16 : (s <= 32) ? 32 : (s <= 48) ? 48 : (s <= 64) ? 64 : (s <= 128) ? 128 : (s + 63) & ~size_t(63);
}

int main(int argc, char** argv)
{
	const int device = argc > 1 ? atoi(argv[1]) : Renderer::addressSpaceOnly;
	const size_t maxN = argc > 2 ? size_t(atoi(argv[2])) : 1030;
	try {
		Renderer r(device);
		DataStorage& ds = r.dataStorage();
		verifyDataStorageEmpty(ds);

		// single allocation of size 10: at the beginning of the buffer, also after free + re-alloc (:58-79)
		HandlelessAllocation a(ds);
		a.alloc(10);
		DataMemory& m = *ds.dataMemoryList().front();
		if(m.size() != Renderer::smallMemorySize) throw std::runtime_error("first DataMemory is not 64 KiB");
		const uint64_t firstAddress = a.deviceAddress();
		if(firstAddress != m.deviceAddress()) throw std::runtime_error("Allocation is not on the beginning of the buffer");
		a.free(); a.free();
		r.executeCopyOperations();
		verifyDataStorageEmpty(ds);
		a.alloc(10);
		if(a.deviceAddress() != m.deviceAddress()) throw std::runtime_error("Allocation is not on the beginning of the buffer");
		r.executeCopyOperations();
		a.free(); a.free();
		verifyDataStorageEmpty(ds);

		// zero-size allocation (:81-86)
		a.alloc(0);
		if(a.deviceAddress() != 0) throw std::runtime_error("Allocation's device address for zero sized allocation is not 0.");
		a.free();
		verifyDataStorageEmpty(ds);

		// single allocation of size 0..260 (:88-96)
		for(size_t s = 0; s < 260; s++) {
			a.alloc(s);
			if(a.deviceAddress() != firstAddress && (a.deviceAddress() != 0 || s != 0)) throw std::runtime_error("Allocation is not on the beginning of the buffer");
			a.free();
			r.executeCopyOperations();
			verifyDataStorageEmpty(ds);
		}

		// two allocations of size 0..260, both release orders (:98-130)
		for(int order = 0; order < 2; order++)
			for(size_t s = 0; s < 260; s++) {
				HandlelessAllocation a1(ds), a2(ds);
				a1.alloc(s); a2.alloc(s);
				if(a1.deviceAddress() != firstAddress && (a1.deviceAddress() != 0 || s != 0)) throw std::runtime_error("Allocation is not on the beginning of the buffer");
				if(a2.deviceAddress() != firstAddress + slot(s) && (a2.deviceAddress() != 0 || s != 0)) throw std::runtime_error("Allocation is not on the the proper place in the buffer");
				if(order == 0) { a1.free(); a2.free(); } else { a2.free(); a1.free(); }
				r.executeCopyOperations();
				verifyDataStorageEmpty(ds);
			}

		// 0..1030 allocations of size 0..260 within 64 KiB, eight release orders (:196-323)
		const size_t sizes[] = {0, 1, 16, 17, 32, 48, 64, 65, 100, 128, 129, 200, 259};
		size_t cases = 0;
		for(size_t s : sizes) {
			const size_t offset = slot(s);
			for(size_t n = 0; n < maxN; n = (n < 8) ? n + 1 : (n < 190 ? n + 37 : (n < 210 ? n + 1 : n + 101))) {
				if(offset * n >= 65536) continue;
				std::vector<HandlelessAllocation> v;
				v.reserve(1030);
				auto fill = [&]() { for(size_t i = 0; i < n; i++) v.emplace_back(ds).alloc(s); };
				auto done = [&]() { v.clear(); r.executeCopyOperations(); verifyDataStorageEmpty(ds); cases++; };
				fill();
				for(size_t i = 0; i < n; i++)
					if(v[i].deviceAddress() != firstAddress + offset * i && (v[i].deviceAddress() != 0 || s != 0))
						throw std::runtime_error("Allocation is not on the the proper place in the buffer (size " + std::to_string(s) + ", n " + std::to_string(n) + ", i " + std::to_string(i) + ")");
				for(size_t i = 0; i < n; i++) v[i].free();
				done();
				fill(); for(size_t i = n; i > 0;) v[--i].free(); done();
				fill(); for(size_t i = 0; i < n; i += 2) v[i].free(); for(size_t i = 1; i < n; i += 2) v[i].free(); done();
				fill(); for(size_t i = 1; i < n; i += 2) v[i].free(); for(size_t i = 0; i < n; i += 2) v[i].free(); done();
				fill(); for(int64_t i = int64_t(n) - 2; i >= 0; i -= 2) v[size_t(i)].free(); for(int64_t i = int64_t(n) - 1; i >= 0; i -= 2) v[size_t(i)].free(); done();
				fill(); for(int64_t i = int64_t(n) - 1; i >= 0; i -= 2) v[size_t(i)].free(); for(int64_t i = int64_t(n) - 2; i >= 0; i -= 2) v[size_t(i)].free(); done();
				// allocations that wrap into region 2 behind a big block (:283-321)
				for(int order = 0; order < 2; order++) {
					HandlelessAllocation bigAllocation(ds);
					bigAllocation.alloc(65536 - s);
					if(n != 0) v.emplace_back(ds).alloc(s);
					bigAllocation.free();
					for(size_t i = 1; i < n; i++) v.emplace_back(ds).alloc(s);
					if(order == 0) for(size_t i = 0; i < n; i++) v[i].free();
					else for(size_t i = n; i > 0;) v[--i].free();
					done();
				}
			}
		}
		printf("data_allocation_test ok (%zu grid cases, %zu DataMemory objects)\n", cases, ds.dataMemoryList().size());
	}
	catch(std::exception& e) { fprintf(stderr, "FAILED: %s\n", e.what()); return 1; }
	catch(Error& e) { fprintf(stderr, "FAILED: %s\n", e.what()); return 1; }
	return 0;
}
