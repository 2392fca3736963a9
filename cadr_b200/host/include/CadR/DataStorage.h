// DataStorage — pool of DataMemory buffers, the handle table and the upload recorder.
// Reference: src/CadR/DataStorage.{h,cpp}.
#pragma once
#include <CadR/DataMemory.h>
#include <CadR/HandleTable.h>
#include <functional>
#include <tuple>
#include <vector>

namespace CadR {

class Renderer;

/// Releases staging resources after the transfer has completed (reference: TransferResources.h).
class TransferResources {
	std::function<void()> _release;
public:
	TransferResources() = default;
	explicit TransferResources(std::function<void()> f) : _release(std::move(f)) {}
	TransferResources(TransferResources&&) = default;
	TransferResources& operator=(TransferResources&&) = default;
	~TransferResources() { release(); }
	void release() { if(_release) { auto f = std::move(_release); _release = nullptr; f(); } }
};

class DataStorage {
	friend class DataMemory;
	friend class HandlelessAllocation;
	friend class DataAllocation;
	Renderer* _renderer;
	std::vector<DataMemory*> _dataMemoryList;
	DataMemory* _firstAllocMemory = nullptr;
	DataMemory* _secondAllocMemory = nullptr;
	DataMemory* _zeroSizeDataMemory;
	DataAllocationRecord _zeroSizeAllocationRecord;
	StagingManager* _stagingManager = nullptr;
	size_t _stagingDataSizeHint = 0;
	uint64_t _uploadEpoch = 1;             // bumped by every recordUploads() that transferred something
	bool _reuseEmptyMemories = true;
	size_t _reuseScanStart = 0;
	HandleTable _handleTable;
	DataAllocationRecord* allocInternal(size_t numBytes);
	std::tuple<StagingMemory&, bool> allocStagingMemory(DataMemory& m, StagingMemory* lastStagingMemory,
	                                                    size_t minNumBytes, size_t bytesToMemoryEnd);
public:
	explicit DataStorage(Renderer& r);
	~DataStorage() noexcept;
	void init(StagingManager& stagingManager) { _stagingManager = &stagingManager; }
	void cleanUp() noexcept;
	DataStorage(const DataStorage&) = delete;

	std::vector<DataMemory*>& dataMemoryList() { return _dataMemoryList; }
	const std::vector<DataMemory*>& dataMemoryList() const { return _dataMemoryList; }
	Renderer& renderer() const { return *_renderer; }
	StagingManager& stagingManager() const { return *_stagingManager; }
	size_t stagingDataSizeHint() const { return _stagingDataSizeHint; }
	uint64_t uploadEpoch() const { return _uploadEpoch; }
	/// Beyond the reference (on by default): when the first and second DataMemory are full, an older DataMemory that has
	/// become completely empty is taken into service again before a new one is created.  The reference only ever moves on
	/// to a NEW DataMemory (DataStorage.cpp:72-85) and never returns to or releases an old one, so a scene that keeps
	/// re-writing its allocations (realloc-on-write) grows by the re-written bytes every frame without bound.  Placement
	/// inside a DataMemory is unaffected; off = the reference's policy.
	void setReuseEmptyDataMemories(bool on) { _reuseEmptyMemories = on; }
	bool reuseEmptyDataMemories() const { return _reuseEmptyMemories; }
	/// Was this record staged in the current frame AND has nothing been transferred since (its staging block is still its own)?
	bool stagedAndNotYetTransferred(const DataAllocationRecord* a) const;
	void setStagingDataSizeHint(size_t size) { _stagingDataSizeHint = size; }

	DataAllocationRecord* alloc(size_t numBytes);
	DataAllocationRecord* realloc(DataAllocationRecord* allocationRecord, size_t numBytes);
	DataAllocationRecord* zeroSizeAllocationRecord() noexcept { return &_zeroSizeAllocationRecord; }
	void free(DataAllocationRecord* a) noexcept { if(a->size == 0) return; DataMemory::free(a); }
	/// DataStorage.cpp:168-172: forwards to every DataMemory, where the reference's implementation is compiled out.
	void cancelAllAllocations() noexcept { for(DataMemory* m : _dataMemoryList) m->cancelAllAllocations(); }

	/// One cadr_b200_upload per call: -> {resources to release once the stream has passed the copies, bytes}.
	std::tuple<TransferResources, size_t> recordUploads(void* stream);
	/// Tools/tests: sees every batch of copy regions {device address, absolute host source, bytes} before it is issued.
	std::function<void(const cadr_copy_region*, size_t)> uploadObserver;

	uint64_t createHandle() { return _handleTable.create(); }
	void destroyHandle(uint64_t handle) noexcept { _handleTable.destroy(handle); }
	void setHandle(uint64_t handle, uint64_t addr) { _handleTable.set(handle, addr); }
	unsigned handleLevel() const { return _handleTable.handleLevel(); }
	uint64_t handleTableDeviceAddress() const { return _handleTable.rootTableDeviceAddress(); }
	const HandleTable& handleTable() const { return _handleTable; }
};

}
