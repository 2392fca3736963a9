// DataMemory — one device buffer ("arena") with a RingSuballocator and the bookkeeping that turns staged
// allocations into copy regions.  Reference: src/CadR/DataMemory.{h,cpp}.
//
// Upload runs replace the reference's marker records (DataMemory.cpp:237-397): a run is a contiguous stretch of
// allocations of one ring region staged in one StagingMemory block; it becomes exactly one copy region
// {dst = first device address, src = staging start, size = staging end - staging start} (:417-425).  While a run
// is pending it pins its ring region (a zero-size ring record), so the region cannot be reset underneath it.
#pragma once
#include <CadR/DataAllocation.h>
#include <CadR/StagingMemory.h>
#include <vector>

extern "C" { struct cadr_copy_region; }

namespace CadR {

class DataMemory {
	friend class DataStorage;
	DataStorage* _dataStorage;
	uint64_t _bufferStart = 0;
	size_t _size = 0;
	bool _ownsBuffer = false;
	RingSuballocator<DataAllocationRecord> _ring;
	struct Run {
		uint64_t deviceAddress;
		StagingMemory* staging;
		uint64_t stagingStart, stagingEnd;
		DataAllocationRecord* pin;
		int slot;                         // RingSuballocator::physicalIndex of the region the run lies in
	};
	// Pending runs and the last staging block per PHYSICAL ring region (RingSuballocator::physicalIndex), in creation
	// order.  Keyed by identity, not by the role "region 1 | 2": the ring swaps the roles when region 1 empties, and a
	// pending run has to stay with the addresses it covers.
	std::vector<Run> _runs[2];
	StagingMemory* _lastStaging[2] = {nullptr, nullptr};
	Run& newRun(int region, uint64_t addr, size_t numBytes, StagingMemory* previous);
public:
	struct PendingUpload { std::vector<Run> runs; };

	DataMemory(DataStorage& storage, size_t size);                     ///< allocates a device buffer (throws OutOfResources)
	DataMemory(DataStorage& storage, uint64_t bufferAddress, size_t size);  ///< adopts an address range (not owned)
	~DataMemory();
	static DataMemory* tryCreate(DataStorage& storage, size_t size);   ///< nullptr instead of throwing (DataMemory.cpp:135-196)

	DataAllocationRecord* alloc(size_t numBytes);                      ///< nullptr when there is no room (DataMemory.cpp:255-397)
	static void free(DataAllocationRecord* a) noexcept;
	/// Present for API compatibility and, as in the reference, does nothing: the body of DataMemory::cancelAllAllocations
	/// is compiled out there ("disabled as it is questionable what it should exactly do", DataMemory.cpp:199-201).
	void cancelAllAllocations() noexcept {}

	/// Append this buffer's copy regions; -> bytes to transfer.  The runs move into `pending` until uploadDone().
	size_t recordUploads(std::vector<cadr_copy_region>& regions, PendingUpload& pending);
	void uploadDone(PendingUpload& pending) noexcept;                  ///< DataMemory.cpp:449-509

	uint64_t deviceAddress() const { return _bufferStart; }
	size_t size() const { return _size; }
	size_t usedBytes() const { return _ring.usedBytes(); }
	bool ringEmpty() const { return _ring.empty(); }
	DataStorage& dataStorage() const { return *_dataStorage; }
};

}
