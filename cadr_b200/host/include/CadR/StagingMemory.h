// StagingMemory / StagingManager — pinned host blocks that mirror a stretch of a DataMemory buffer.
// Reference: src/CadR/StagingMemory.{h,cpp} (HOST_VISIBLE|HOST_CACHED mapped buffers, :55-75) and
// src/CadR/StagingManager.{h,cpp} (size-class pools with reuse lists, :27-86).  Here a block is pinned host
// memory from cadr_b200_host_alloc, so every upload is a true DMA.
#pragma once
#include <cstddef>
#include <cstdint>
#include <list>
#include <memory>

struct cadr_ctx;

namespace CadR {

class StagingManager;

class StagingMemory {
	friend class StagingManager;
	friend class DataMemory;
	StagingManager* _manager;
	uint8_t* _host = nullptr;
	size_t _size = 0;
	int64_t _deviceToStaging = 0;   ///< stagingAddress = deviceAddress + _deviceToStaging   (StagingMemory.h:38)
	uint64_t _referenceCounter = 0;
public:
	StagingMemory(StagingManager& manager, size_t size);
	~StagingMemory();
	StagingMemory(const StagingMemory&) = delete;
	template<typename T = void> T* data() { return reinterpret_cast<T*>(_host); }
	size_t size() const { return _size; }
	uint64_t hostStart() const { return reinterpret_cast<uint64_t>(_host); }
	uint64_t hostEnd() const { return hostStart() + _size; }
	/// Does the device range [addr, addr+bytes) fall outside this block under the current mapping?  Two-sided: a device
	/// address below the start of the stretch a non-exclusive block mirrors maps in FRONT of the block.
	bool addrRangeOverruns(uint64_t deviceAddr, size_t bytes) const {
		const uint64_t s = uint64_t(int64_t(deviceAddr) + _deviceToStaging);
		return s < hostStart() || s + bytes > hostEnd();
	}
};

class StagingManager {
	cadr_ctx* _ctx;
	using List = std::list<std::unique_ptr<StagingMemory>>;
	List _inUse[4], _available[4];   // small, medium, large, super-size
	static int classOf(size_t size);
	StagingMemory& reuseOrAlloc(int cls, size_t size);
public:
	static constexpr size_t smallMemorySize = 64 << 10;    // Renderer.h:104-106
	static constexpr size_t mediumMemorySize = 2 << 20;
	static constexpr size_t largeMemorySize = 32 << 20;
	explicit StagingManager(cadr_ctx* ctx) : _ctx(ctx) {}
	~StagingManager() { cleanUp(); }
	void cleanUp() noexcept;
	cadr_ctx* context() const { return _ctx; }
	StagingMemory& reuseOrAllocSmallStagingMemory() { return reuseOrAlloc(0, smallMemorySize); }
	StagingMemory& reuseOrAllocMediumStagingMemory() { return reuseOrAlloc(1, mediumMemorySize); }
	StagingMemory& reuseOrAllocLargeStagingMemory() { return reuseOrAlloc(2, largeMemorySize); }
	StagingMemory& reuseOrAllocSuperSizeStagingMemory(size_t size);
	void freeOrRecycleStagingMemory(StagingMemory& sm) noexcept;
	size_t numBlocksInUse() const { size_t n = 0; for(auto& l : _inUse) n += l.size(); return n; }
	size_t numBlocksAvailable() const { size_t n = 0; for(auto& l : _available) n += l.size(); return n; }
};

}
