// Memory layer of the facade: staging pools, DataMemory, DataStorage, allocations, handle table.
// Semantics follow the reference files cited in the headers; the code is written against the C ABI
// (include/cadr_b200.h) and shares no structure with the Vulkan implementation.
#include <CadR/CadR.h>
#include "../../../include/cadr_b200.h"
#include <algorithm>
#include <cstring>

namespace CadR {

void check(int code)
{
	if(code == CADR_OK) return;
	std::string msg = cadr_b200_last_error();
	switch(code) {
	case CADR_E_LOGIC: throw LogicError(msg);
	case CADR_E_OUT_OF_RESOURCES: throw OutOfResources(msg);
	case CADR_E_TIMEOUT: throw Timeout(msg);
	default: throw DeviceError(msg);
	}
}

// ---------------------------------------------------------------------------------------------------
// StagingMemory / StagingManager
// ---------------------------------------------------------------------------------------------------
StagingMemory::StagingMemory(StagingManager& manager, size_t size) : _manager(&manager), _size(size)
{
	void* p = nullptr;
	check(cadr_b200_host_alloc(manager.context(), size, &p));
	_host = static_cast<uint8_t*>(p);
}

StagingMemory::~StagingMemory()
{
	if(_host) cadr_b200_host_free(_manager->context(), _host);
}

int StagingManager::classOf(size_t size)
{
	return size <= smallMemorySize ? 0 : size <= mediumMemorySize ? 1 : size <= largeMemorySize ? 2 : 3;
}

StagingMemory& StagingManager::reuseOrAlloc(int cls, size_t size)
{
	if(!_available[cls].empty()) {
		_inUse[cls].splice(_inUse[cls].end(), _available[cls], _available[cls].begin());
		return *_inUse[cls].back();
	}
	_inUse[cls].push_back(std::make_unique<StagingMemory>(*this, size));
	return *_inUse[cls].back();
}

StagingMemory& StagingManager::reuseOrAllocSuperSizeStagingMemory(size_t size)
{
	// smallest available block that is large enough (StagingManager.cpp:41-62)
	auto best = _available[3].end();
	for(auto it = _available[3].begin(); it != _available[3].end(); ++it)
		if((*it)->size() >= size && (best == _available[3].end() || (*it)->size() < (*best)->size()))
			best = it;
	if(best != _available[3].end()) {
		_inUse[3].splice(_inUse[3].end(), _available[3], best);
		return *_inUse[3].back();
	}
	_inUse[3].push_back(std::make_unique<StagingMemory>(*this, size));
	return *_inUse[3].back();
}

void StagingManager::freeOrRecycleStagingMemory(StagingMemory& sm) noexcept
{
	int cls = classOf(sm.size());
	for(auto it = _inUse[cls].begin(); it != _inUse[cls].end(); ++it)
		if(it->get() == &sm) {
			_available[cls].splice(_available[cls].begin(), _inUse[cls], it);
			return;
		}
}

void StagingManager::cleanUp() noexcept
{
	for(auto& l : _inUse) l.clear();
	for(auto& l : _available) l.clear();
}

// ---------------------------------------------------------------------------------------------------
// DataMemory
// ---------------------------------------------------------------------------------------------------
DataMemory::DataMemory(DataStorage& storage, size_t size) : _dataStorage(&storage), _size(size)
{
	if(size == 0) { _ring.reset(0, 0); return; }
	check(cadr_b200_arena_alloc(storage.stagingManager().context(), size, &_bufferStart));
	_ownsBuffer = true;
	_ring.reset(_bufferStart, size);
}

DataMemory::DataMemory(DataStorage& storage, uint64_t bufferAddress, size_t size)
	: _dataStorage(&storage), _bufferStart(bufferAddress), _size(size)
{
	_ring.reset(bufferAddress, size);
}

DataMemory::~DataMemory()
{
	// pending runs still hold staging blocks and pins (DataMemory.cpp:17-25)
	for(auto& runs : _runs) {
		PendingUpload p{std::move(runs)};
		uploadDone(p);
	}
	if(_ownsBuffer) cadr_b200_arena_free(_dataStorage->stagingManager().context(), _bufferStart);
}

DataMemory* DataMemory::tryCreate(DataStorage& storage, size_t size)
{
	try { return new DataMemory(storage, size); }
	catch(OutOfResources&) { return nullptr; }
}

DataMemory::Run& DataMemory::newRun(int region, uint64_t addr, size_t numBytes, StagingMemory* previous)
{
	const int idx = _ring.physicalIndex(region);
	StagingMemory* other = _lastStaging[1 - idx];
	StagingMemory* sm;
	uint64_t stagingStart;
	// the pin goes in first: it precedes the run's first allocation in the ring, like the reference's marker
	DataAllocationRecord* pin = _ring.pin(region, addr);
	try {
		if(other && other->size() == _size) {
			// an exclusive block mirrors the whole buffer: both regions can share it (DataMemory.cpp:312-320)
			sm = other;
			stagingStart = uint64_t(int64_t(addr) + sm->_deviceToStaging);
		}
		else {
			auto [block, exclusive] = _dataStorage->allocStagingMemory(*this, previous, numBytes, size_t(_bufferStart + _size - addr));
			sm = &block;
			if(exclusive) {
				sm->_deviceToStaging = int64_t(sm->hostStart()) - int64_t(_bufferStart);
				stagingStart = uint64_t(int64_t(addr) + sm->_deviceToStaging);
			}
			else {
				sm->_deviceToStaging = int64_t(sm->hostStart()) - int64_t(addr);
				stagingStart = sm->hostStart();
			}
		}
	}
	catch(...) { _ring.release(pin); throw; }
	sm->_referenceCounter++;
	_lastStaging[idx] = sm;
	_runs[idx].push_back(Run{addr, sm, stagingStart, stagingStart, pin, idx});
	return _runs[idx].back();
}

DataAllocationRecord* DataMemory::alloc(size_t numBytes)
{
	auto [addr, region] = _ring.propose(numBytes);
	if(region == 0) return nullptr;
	// runs and staging blocks belong to the physical region (the ring may have swapped the roles of its two regions since
	// the pending run was opened: it stays with its addresses)
	const int idx = _ring.physicalIndex(region);
	StagingMemory* last = _lastStaging[idx];
	Run* run = _runs[idx].empty() ? nullptr : &_runs[idx].back();
	if(run) {
		// a run is ONE copy region: it is continued only by an allocation that lies behind everything it holds
		// (DataMemory.cpp:417-425 computes size = stagingEnd - stagingStart), inside the same staging block
		const bool extends = run->staging == last && !last->addrRangeOverruns(addr, numBytes) &&
		                     uint64_t(int64_t(addr) + last->_deviceToStaging) >= run->stagingEnd;
		if(!extends)
			run = &newRun(region, addr, numBytes, last->addrRangeOverruns(addr, numBytes) ? last : nullptr);   // block full: a bigger one
	}
	else if(last && !last->addrRangeOverruns(addr, numBytes)) {
		// first allocation since the last transfer and the previous block still has room behind it
		DataAllocationRecord* pin = _ring.pin(region, addr);
		last->_referenceCounter++;
		uint64_t s = uint64_t(int64_t(addr) + last->_deviceToStaging);
		_runs[idx].push_back(Run{addr, last, s, s, pin, idx});
		run = &_runs[idx].back();
	}
	else
		run = &newRun(region, addr, numBytes, last);

	DataAllocationRecord* a = _ring.commit(region, addr, numBytes);
	const uint64_t stagingAddr = uint64_t(int64_t(addr) + run->staging->_deviceToStaging);
	a->dataMemory = this;
	a->recordPointer = nullptr;
	a->stagingData = reinterpret_cast<void*>(stagingAddr);
	a->stagingFrameNumber = size_t(-2);
	run->stagingEnd = stagingAddr + numBytes;
	return a;
}

void DataMemory::free(DataAllocationRecord* a) noexcept
{
	a->dataMemory->_ring.release(a);
}

size_t DataMemory::recordUploads(std::vector<cadr_copy_region>& regions, PendingUpload& pending)
{
	size_t bytes = 0;
	for(auto& runs : _runs) {
		for(Run& r : runs) {
			const uint64_t size = r.stagingEnd - r.stagingStart;
			regions.push_back(cadr_copy_region{r.deviceAddress, r.stagingStart, size});   // src is an absolute host address
			bytes += size;
			pending.runs.push_back(r);
		}
		runs.clear();
	}
	return bytes;
}

void DataMemory::uploadDone(PendingUpload& pending) noexcept
{
	for(Run& r : pending.runs) {
		StagingMemory* sm = r.staging;
		if(--sm->_referenceCounter == 0) {
			if(_lastStaging[0] == sm) _lastStaging[0] = nullptr;
			if(_lastStaging[1] == sm) _lastStaging[1] = nullptr;
			_dataStorage->stagingManager().freeOrRecycleStagingMemory(*sm);
		}
		_ring.release(r.pin);
	}
	pending.runs.clear();
}

// ---------------------------------------------------------------------------------------------------
// DataStorage
// ---------------------------------------------------------------------------------------------------
DataStorage::DataStorage(Renderer& r) : _renderer(&r), _handleTable(*this)
{
	_zeroSizeDataMemory = nullptr;
	_zeroSizeAllocationRecord.deviceAddress = 0;
	_zeroSizeAllocationRecord.size = 0;
}

DataStorage::~DataStorage() noexcept { cleanUp(); }

void DataStorage::cleanUp() noexcept
{
	_handleTable.destroyAll();
	for(DataMemory* m : _dataMemoryList) delete m;
	_dataMemoryList.clear();
	_firstAllocMemory = _secondAllocMemory = nullptr;
}

DataAllocationRecord* DataStorage::allocInternal(size_t numBytes)
{
	using R = Renderer;
	auto create = [&](size_t size) {
		DataMemory* m = DataMemory::tryCreate(*this, size);
		if(!m) throw OutOfResources("CadR::DataStorage::alloc() error: Cannot allocate DataMemory. Requested size: " + std::to_string(size) + " bytes.");
		_dataMemoryList.push_back(m);
		return m;
	};
	// three attempts: first buffer, second buffer, a fresh large buffer (DataStorage.cpp:31-96)
	if(!_firstAllocMemory)
		_firstAllocMemory = create(numBytes < R::smallMemorySize ? R::smallMemorySize
		                           : numBytes < R::mediumMemorySize ? R::mediumMemorySize : std::max(numBytes, R::largeMemorySize));
	if(DataAllocationRecord* a = _firstAllocMemory->alloc(numBytes)) return a;
	if(!_secondAllocMemory)
		_secondAllocMemory = create(numBytes < R::mediumMemorySize ? R::mediumMemorySize : std::max(numBytes, R::largeMemorySize));
	if(DataAllocationRecord* a = _secondAllocMemory->alloc(numBytes)) return a;
	// (extension, see DataStorage::setReuseEmptyDataMemories) an old DataMemory that has emptied out completely - no live
	// allocation, no pending upload run - serves again before a new device buffer is created; the scan resumes where the
	// last one ended, so that a long list of full memories is not walked from its start every time
	DataMemory* m = nullptr;
	if(_reuseEmptyMemories) {
		const size_t count = _dataMemoryList.size();
		for(size_t i = 0; i < count && !m; i++) {
			DataMemory* c = _dataMemoryList[(_reuseScanStart + i) % count];
			if(c != _firstAllocMemory && c != _secondAllocMemory && c->size() >= std::max(R::largeMemorySize, numBytes) && c->ringEmpty()) {
				m = c;
				_reuseScanStart = (_reuseScanStart + i + 1) % count;
			}
		}
	}
	if(!m) m = create(std::max(R::largeMemorySize, numBytes));
	_firstAllocMemory = _secondAllocMemory;   // the first is full, the second nearly: rotate
	_secondAllocMemory = m;
	DataAllocationRecord* a = m->alloc(numBytes);
	if(!a) throw OutOfResources("CadR::DataStorage::alloc() error: Cannot allocate DataAllocation although new DataMemory was created successfully.");
	return a;
}

DataAllocationRecord* DataStorage::alloc(size_t numBytes)
{
	if(numBytes == 0) return &_zeroSizeAllocationRecord;   // null-object pattern (DataStorage.cpp:121-127)
	DataAllocationRecord* a = allocInternal(numBytes);
	a->stagingFrameNumber = _renderer->frameNumber();
	a->stagingEpoch = _uploadEpoch;
	return a;
}

bool DataStorage::stagedAndNotYetTransferred(const DataAllocationRecord* a) const
{
	return a->stagingFrameNumber == _renderer->frameNumber() && a->stagingEpoch == _uploadEpoch;
}

DataAllocationRecord* DataStorage::realloc(DataAllocationRecord* allocationRecord, size_t numBytes)
{
	if(numBytes == 0) { free(allocationRecord); return &_zeroSizeAllocationRecord; }
	DataAllocationRecord* a = allocInternal(numBytes);   // throws before anything is released
	a->stagingFrameNumber = _renderer->frameNumber();
	a->stagingEpoch = _uploadEpoch;
	if(allocationRecord->size != 0) DataMemory::free(allocationRecord);
	return a;
}

std::tuple<StagingMemory&, bool> DataStorage::allocStagingMemory(DataMemory& m, StagingMemory* last, size_t minNumBytes, size_t bytesToMemoryEnd)
{
	using R = Renderer;
	StagingManager& sm = *_stagingManager;
	if(last) {
		// the previous block filled up: step up one size class (DataStorage.cpp:180-196)
		if(m.size() <= R::mediumMemorySize) return {sm.reuseOrAllocMediumStagingMemory(), true};
		if(last->size() < R::mediumMemorySize) return {sm.reuseOrAllocMediumStagingMemory(), false};
		if(m.size() <= R::largeMemorySize) return {sm.reuseOrAllocLargeStagingMemory(), true};
		if(last->size() < R::largeMemorySize) return {sm.reuseOrAllocLargeStagingMemory(), false};
		return {sm.reuseOrAllocSuperSizeStagingMemory(m.size()), true};
	}
	const float hint = float(_stagingDataSizeHint) * 1.2f;
	if(m.size() <= R::smallMemorySize) return {sm.reuseOrAllocSmallStagingMemory(), true};
	if(bytesToMemoryEnd <= R::smallMemorySize || (minNumBytes <= R::smallMemorySize && hint < R::smallMemorySize))
		return {sm.reuseOrAllocSmallStagingMemory(), false};
	if(m.size() <= R::mediumMemorySize) return {sm.reuseOrAllocMediumStagingMemory(), true};
	if(bytesToMemoryEnd <= R::mediumMemorySize || (minNumBytes <= R::mediumMemorySize && hint < R::mediumMemorySize))
		return {sm.reuseOrAllocMediumStagingMemory(), false};
	if(m.size() <= R::largeMemorySize) return {sm.reuseOrAllocLargeStagingMemory(), true};
	if(bytesToMemoryEnd <= R::largeMemorySize || (minNumBytes <= R::largeMemorySize && hint < R::largeMemorySize))
		return {sm.reuseOrAllocLargeStagingMemory(), false};
	return {sm.reuseOrAllocSuperSizeStagingMemory(minNumBytes), true};
}

std::tuple<TransferResources, size_t> DataStorage::recordUploads(void* stream)
{
	std::vector<cadr_copy_region> regions;
	auto pending = std::make_shared<std::vector<std::pair<DataMemory*, DataMemory::PendingUpload>>>();
	size_t bytes = 0;
	for(DataMemory* dm : _dataMemoryList) {
		DataMemory::PendingUpload p;
		size_t b = dm->recordUploads(regions, p);
		if(!p.runs.empty()) pending->emplace_back(dm, std::move(p));
		bytes += b;
	}
	if(pending->empty()) return {TransferResources(), 0};
	_uploadEpoch++;          // staging blocks of everything staged so far are released when this transfer completes
	TransferResources tr([pending]() { for(auto& [dm, p] : *pending) dm->uploadDone(p); });
	if(uploadObserver) uploadObserver(regions.data(), regions.size());
	cadr_ctx* ctx = _stagingManager->context();
	if(bytes != 0 && cadr_b200_device(ctx) >= 0)
		check(cadr_b200_upload(ctx, regions.data(), uint32_t(regions.size()), nullptr, stream));
	return {std::move(tr), bytes};
}

// ---------------------------------------------------------------------------------------------------
// HandlelessAllocation / DataAllocation
// ---------------------------------------------------------------------------------------------------
HandlelessAllocation::HandlelessAllocation(DataStorage& storage) noexcept : _record(storage.zeroSizeAllocationRecord()), _storage(&storage) {}
HandlelessAllocation::HandlelessAllocation(Renderer& r) noexcept : HandlelessAllocation(r.dataStorage()) {}
HandlelessAllocation::HandlelessAllocation(HandlelessAllocation&& o) noexcept : _record(o._record), _storage(o._storage)
{
	o._record = _storage->zeroSizeAllocationRecord();
}
HandlelessAllocation& HandlelessAllocation::operator=(HandlelessAllocation&& rhs) noexcept
{
	if(this != &rhs) { free(); _record = rhs._record; _storage = rhs._storage; rhs._record = _storage->zeroSizeAllocationRecord(); }
	return *this;
}
Renderer& HandlelessAllocation::renderer() const { return _storage->renderer(); }
size_t HandlelessAllocation::offset() const { return size_t(_record->deviceAddress - _record->dataMemory->deviceAddress()); }

StagingData HandlelessAllocation::alloc(size_t size)
{
	// reuse what was staged earlier in this frame (DataAllocation.cpp:54-69)
	if(_storage->stagedAndNotYetTransferred(_record) && size <= _record->size) {
		_record->size = size;
		return StagingData(_record, false);
	}
	_record = _storage->realloc(_record, size);
	return StagingData(_record, true);
}

StagingData HandlelessAllocation::alloc()
{
	if(_storage->stagedAndNotYetTransferred(_record))
		return StagingData(_record, false);
	_record = _storage->realloc(_record, _record->size);
	return StagingData(_record, true);
}

void HandlelessAllocation::free() noexcept
{
	if(_record->size == 0) return;
	DataMemory::free(_record);
	_record = _storage->zeroSizeAllocationRecord();
}

void HandlelessAllocation::upload(const void* ptr, size_t numBytes)
{
	_record = _storage->realloc(_record, numBytes);
	std::memcpy(_record->stagingData, ptr, numBytes);
}

DataAllocation::DataAllocation(DataStorage& storage) : HandlelessAllocation(storage), _handle(storage.createHandle()) {}
DataAllocation::DataAllocation(DataStorage& storage, noHandle_t) noexcept : HandlelessAllocation(storage), _handle(0) {}
DataAllocation::DataAllocation(Renderer& r) : DataAllocation(r.dataStorage()) {}
DataAllocation::DataAllocation(Renderer& r, noHandle_t) noexcept : DataAllocation(r.dataStorage(), noHandle) {}
DataAllocation::DataAllocation(DataAllocation&& o) noexcept : HandlelessAllocation(std::move(o)), _handle(o._handle) { o._handle = 0; }
DataAllocation::~DataAllocation() noexcept { free(); if(_handle != 0) _storage->destroyHandle(_handle); }
DataAllocation& DataAllocation::operator=(DataAllocation&& rhs) noexcept
{
	if(this != &rhs) {
		free();
		if(_handle != 0) _storage->destroyHandle(_handle);
		HandlelessAllocation::operator=(std::move(rhs));
		_handle = rhs._handle; rhs._handle = 0;
	}
	return *this;
}

StagingData DataAllocation::alloc(size_t numBytes)
{
	StagingData sd = HandlelessAllocation::alloc(numBytes);
	// the handle follows the allocation to its new address (DataAllocation.cpp:31-33); unlike the reference a
	// handle-less allocation does not write table slot 0, which must stay zero for "no drawable data"
	if(sd.wasReallocated() && _handle != 0) _storage->setHandle(_handle, _record->deviceAddress);
	return sd;
}

StagingData DataAllocation::alloc()
{
	StagingData sd = HandlelessAllocation::alloc();
	if(sd.wasReallocated() && _handle != 0) _storage->setHandle(_handle, _record->deviceAddress);
	return sd;
}

void DataAllocation::upload(const void* ptr, size_t numBytes)
{
	HandlelessAllocation::upload(ptr, numBytes);
	if(_handle != 0) _storage->setHandle(_handle, _record->deviceAddress);
}

uint64_t DataAllocation::createHandle(DataStorage& storage) { if(_handle == 0) _handle = storage.createHandle(); return _handle; }
void DataAllocation::destroyHandle() noexcept { if(_handle == 0) return; _storage->destroyHandle(_handle); _handle = 0; }

// ---------------------------------------------------------------------------------------------------
// HandleTable
// ---------------------------------------------------------------------------------------------------
HandleTable::Node::Node(DataStorage& storage, bool routing) : allocation(storage)
{
	if(routing) children = new std::array<std::unique_ptr<Node>, numHandlesPerTable>();
}
HandleTable::Node::~Node() { delete children; }

void HandleTable::Node::init()
{
	constexpr size_t bytes = numHandlesPerTable * sizeof(uint64_t);
	StagingData sd = allocation.alloc(bytes);
	std::memset(sd.data(), 0, bytes);
}

std::unique_ptr<HandleTable::Node> HandleTable::makeNode(bool routing)
{
	auto n = std::make_unique<Node>(*_storage, routing);
	n->init();
	return n;
}

void HandleTable::setEntry(Node& node, unsigned index, uint64_t value)
{
	node.entries[index] = value;
	StagingData sd = node.allocation.createStagingData();
	uint64_t* a = sd.data<uint64_t>();
	if(sd.wasReallocated()) {
		// first write of the frame: the table moved, so the whole node is re-staged and its parent repointed
		std::memcpy(a, node.entries.data(), sizeof(node.entries));
		if(node.parent) setEntry(*node.parent, node.indexInParent, node.allocation.deviceAddress());
	}
	else a[index] = value;
}

uint64_t HandleTable::create()
{
	const uint64_t h = _highestHandle + 1;
	auto adopt = [](Node& parent, unsigned idx, std::unique_ptr<Node> child) -> Node& {
		child->parent = &parent; child->indexInParent = idx;
		(*parent.children)[idx] = std::move(child);
		return *(*parent.children)[idx];
	};
	if(_handleLevel == 0) {
		_root = makeNode(false);
		_handleLevel = 1;
	}
	else if(_handleLevel == 1) {
		if(h == numHandlesPerTable) {
			// one table -> two levels: a routing table over the old leaf and a second leaf (HandleTable.cpp:146-181)
			auto l1 = makeNode(true);
			auto llt = makeNode(false);
			Node& oldLeaf = adopt(*l1, 0, std::move(_root));
			Node& newLeaf = adopt(*l1, 1, std::move(llt));
			_root = std::move(l1);
			setEntry(*_root, 0, oldLeaf.allocation.deviceAddress());
			setEntry(*_root, 1, newLeaf.allocation.deviceAddress());
			_handleLevel = 2;
		}
	}
	else if(_handleLevel == 2) {
		if((h & handleBitsLevelMask) == 0) {
			auto llt = makeNode(false);
			const uint64_t l1Index = h >> handleBitsLevelShift;
			if(l1Index < numHandlesPerTable) {
				Node& leaf = adopt(*_root, unsigned(l1Index), std::move(llt));
				setEntry(*_root, unsigned(l1Index), leaf.allocation.deviceAddress());
			}
			else {
				// two -> three levels (HandleTable.cpp:228-267)
				auto l1 = makeNode(true);
				auto l2 = makeNode(true);
				Node& leaf = adopt(*l1, 0, std::move(llt));
				setEntry(*l1, 0, leaf.allocation.deviceAddress());
				Node& oldL1 = adopt(*l2, 0, std::move(_root));
				Node& newL1 = adopt(*l2, 1, std::move(l1));
				_root = std::move(l2);
				setEntry(*_root, 0, oldL1.allocation.deviceAddress());
				setEntry(*_root, 1, newL1.allocation.deviceAddress());
				_handleLevel = 3;
			}
		}
	}
	else {
		if((h & handleBitsLevelMask) == 0) {
			auto llt = makeNode(false);
			const unsigned l1Index = unsigned(h >> handleBitsLevelShift) & handleBitsLevelMask;
			const unsigned l2Index = unsigned(h >> (2 * handleBitsLevelShift));
			if(l2Index >= numHandlesPerTable) throw OutOfResources("CadR::HandleTable: handle space exhausted");
			if(l1Index != 0) {
				Node& l1 = *(*_root->children)[l2Index];
				Node& leaf = adopt(l1, l1Index, std::move(llt));
				setEntry(l1, l1Index, leaf.allocation.deviceAddress());   // cascades to the root if l1 moved
			}
			else {
				auto l1 = makeNode(true);
				Node& leaf = adopt(*l1, 0, std::move(llt));
				setEntry(*l1, 0, leaf.allocation.deviceAddress());
				Node& nl1 = adopt(*_root, l2Index, std::move(l1));
				setEntry(*_root, l2Index, nl1.allocation.deviceAddress());
			}
		}
	}
	_highestHandle = h;
	return h;
}

HandleTable::Node* HandleTable::leafFor(uint64_t handle) const
{
	switch(_handleLevel) {
	case 1: return _root.get();
	case 2: return (*_root->children)[handle >> handleBitsLevelShift].get();
	case 3: {
		Node* l1 = (*_root->children)[handle >> (2 * handleBitsLevelShift)].get();
		return (*l1->children)[(handle >> handleBitsLevelShift) & handleBitsLevelMask].get();
	}
	default: return nullptr;
	}
}

void HandleTable::set(uint64_t handle, uint64_t addr)
{
	if(_handleLevel == 0) return;                 // HandleTable::setHandle0
	if(handle > _highestHandle) throw LogicError("CadR::HandleTable::set(): handle was never created");
	Node* leaf = leafFor(handle);
	setEntry(*leaf, _handleLevel == 1 ? unsigned(handle) : unsigned(handle & handleBitsLevelMask), addr);
}

uint64_t HandleTable::rootTableDeviceAddress() const { return _root ? _root->allocation.deviceAddress() : 0; }

void HandleTable::destroyAll() noexcept
{
	_root.reset();          // frees every node allocation, children first
	_handleLevel = 0;
	_highestHandle = 0;
}

}
