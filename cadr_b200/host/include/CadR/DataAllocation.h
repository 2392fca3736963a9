// DataAllocation / HandlelessAllocation / StagingData / DataAllocationRecord — one piece of device memory with
// realloc-on-write semantics.  Reference: src/CadR/DataAllocation.{h,cpp}, src/CadR/StagingData.h.
//
//   alloc(n)   in a new frame returns a NEW device range (same handle, new address) whose staging block the caller
//              fills; a second alloc(n' <= size) in the same frame reuses the staged block (DataAllocation.cpp:18-34)
//   handle()   64-bit handle, stable for the object's lifetime; the handle table entry follows the address
#pragma once
#include <CadR/RingSuballocator.h>
#include <cstring>

namespace CadR {

class DataMemory;
class DataStorage;
class Renderer;

struct DataAllocationRecord : RingRecord {     // deviceAddress, size come from RingRecord
	DataMemory* dataMemory = nullptr;
	DataAllocationRecord** recordPointer = nullptr;
	void* stagingData = nullptr;
	size_t stagingFrameNumber = size_t(-2);
	/// DataStorage::uploadEpoch() when the staging block was handed out.  The reference reuses a staged block for every
	/// further write of the same frame (DataAllocation.cpp:22-28, frame number only) - also after executeCopyOperations
	/// has transferred and RELEASED that block in the middle of the frame, when the write lands in recycled staging memory
	/// and never reaches the device.  The facade reuses a block only while nothing has been transferred since.
	uint64_t stagingEpoch = 0;
};

class StagingData {
	DataAllocationRecord* _record = nullptr;
	bool _wasReallocated = false;
public:
	StagingData() = default;
	StagingData(DataAllocationRecord* record, bool wasReallocated) : _record(record), _wasReallocated(wasReallocated) {}
	template<typename T = void> T* data() { return reinterpret_cast<T*>(_record->stagingData); }
	size_t sizeInBytes() const { return _record->size; }
	bool wasReallocated() const { return _wasReallocated; }
};

class HandlelessAllocation {
protected:
	DataAllocationRecord* _record;
	DataStorage* _storage;
public:
	explicit HandlelessAllocation(DataStorage& storage) noexcept;
	explicit HandlelessAllocation(Renderer& r) noexcept;
	HandlelessAllocation(HandlelessAllocation&& other) noexcept;
	HandlelessAllocation(const HandlelessAllocation&) = delete;
	~HandlelessAllocation() noexcept { free(); }
	HandlelessAllocation& operator=(HandlelessAllocation&& rhs) noexcept;
	HandlelessAllocation& operator=(const HandlelessAllocation&) = delete;

	StagingData alloc(size_t size);
	StagingData alloc();
	void free() noexcept;
	StagingData createStagingData() { return alloc(); }
	StagingData createStagingData(size_t size) { return alloc(size); }
	void setData(const void* data, size_t size) { StagingData sd = alloc(size); std::memcpy(sd.data(), data, size); }
	template<typename T> void setData(const T& data) { setData(&data, sizeof(data)); }
	template<typename T> T* editNewContent(size_t count) { StagingData sd = alloc(sizeof(T) * count); return sd.data<T>(); }
	void upload(const void* ptr, size_t numBytes);

	uint64_t deviceAddress() const { return _record->deviceAddress; }
	size_t size() const { return _record->size; }
	size_t offset() const;
	DataMemory& dataMemory() const { return *_record->dataMemory; }
	DataStorage& dataStorage() const { return *_storage; }
	Renderer& renderer() const;
};

class DataAllocation : public HandlelessAllocation {
	uint64_t _handle;
public:
	enum class noHandle_t : int;
	static constexpr noHandle_t noHandle = noHandle_t(0);
	explicit DataAllocation(DataStorage& storage);
	DataAllocation(DataStorage& storage, noHandle_t) noexcept;
	explicit DataAllocation(Renderer& r);
	DataAllocation(Renderer& r, noHandle_t) noexcept;
	DataAllocation(DataAllocation&& other) noexcept;
	~DataAllocation() noexcept;
	DataAllocation& operator=(DataAllocation&& rhs) noexcept;

	StagingData alloc(size_t numBytes);
	StagingData alloc();
	StagingData createStagingData() { return alloc(); }
	StagingData createStagingData(size_t size) { return alloc(size); }
	void setData(const void* data, size_t size) { StagingData sd = alloc(size); std::memcpy(sd.data(), data, size); }
	template<typename T> void setData(const T& data) { setData(&data, sizeof(data)); }
	template<typename T> T* editNewContent(size_t count) { StagingData sd = alloc(sizeof(T) * count); return sd.data<T>(); }
	void upload(const void* ptr, size_t numBytes);
	uint64_t createHandle(DataStorage& storage);
	void destroyHandle() noexcept;
	uint64_t handle() const { return _handle; }
};

}
