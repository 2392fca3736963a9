// Renderer — the frame API of the facade.  Reference: src/CadR/Renderer.{h,cpp}.
//
// Same call order per frame as examples/RenderingPerformance/main.cpp:1424-1572:
//   beginFrame -> edits -> executeCopyOperations -> beginRecording -> prepareSceneRendering ->
//   recordDrawableProcessing -> recordSceneRendering -> endRecording -> executeCopyOperations -> submit -> endFrame
// Divergences, all forced by the absence of Vulkan types (vulkan.hpp does not even compile here):
//   * vk::CommandBuffer parameters are gone: recording appends to an internal command list, and submit() — the
//     stand-in for the application's vkQueueSubmit — enqueues it on the CUDA stream (DMA of the drawable list,
//     processing kernel, optionally the culling kernels);
//   * getters that returned vk::Buffer return the device address;
//   * recordSceneRendering does not bind pipelines or issue vkCmdDrawIndirect: it produces the list of draw
//     ranges {StateSet, first drawable, count, pointers base, indirect offset} a rasteriser would consume.
#pragma once
#include <CadR/DataStorage.h>
#include <CadR/Exceptions.h>
#include <CadR/Scene.h>
#include <CadR/StagingMemory.h>
#include <array>
#include <memory>

struct cadr_ctx;

namespace CadR {

struct FrameInfo {                // src/CadR/FrameInfo.h (GPU timestamps become milliseconds since beginRecording)
	size_t frameNumber = ~size_t(0);
	double cpuBeginFrame = 0, cpuPrepareRecordingBegin = 0, cpuPrepareRecordingEnd = 0;
	double cpuRecordStateSetsBegin = 0, cpuRecordStateSetsEnd = 0, cpuEndFrame = 0;     // seconds, steady clock
	float gpuBeginExecution = 0, gpuAfterTransfersAndBeforeDrawableProcessing = 0;
	float gpuAfterDrawableProcessingAndBeforeRendering = 0, gpuEndExecution = 0;         // milliseconds
	static constexpr uint32_t gpuTimestampPoolSize = 4;
};

/// What StateSet::recordToCommandBuffer would have drawn (StateSet.cpp:240-258).
struct DrawRange {
	StateSet* stateSet;
	size_t firstDrawable;
	size_t numDrawables;
	uint64_t drawablePointersAddress;   ///< push constant @8: pointers base + firstDrawable*32
	uint64_t indirectOffset;            ///< vkCmdDrawIndirect offset: firstDrawable*16
};

struct Frustum {                  // inputs of the culling extension
	float planes[6][4];
	float eye[3];
};

struct CullResult {               // device buffers of the culling extension (include/cadr_b200.h)
	uint64_t commands = 0, pointers = 0, tags = 0, instances = 0, counters = 0;
	uint32_t numRanges = 0;
	std::vector<std::array<uint32_t, 4>> regions;   ///< per draw range: cmdBase, cmdCapacity, instBase, instCapacity
};

class Renderer {
	friend class StateSet;
	cadr_ctx* _ctx = nullptr;
	bool _ownsContext = false;
	void* _stream = nullptr;
	std::unique_ptr<StagingManager> _stagingManager;
	std::unique_ptr<DataStorage> _dataStorage;
	size_t _frameNumber = ~size_t(0);
	size_t _currentFrameUploadBytes = 0, _lastFrameUploadBytes = 0;
	// drawable buffers (Renderer.cpp:461-592)
	size_t _drawableCapacity = 0;
	uint64_t _drawableBufferAddress = 0, _drawIndirectBufferAddress = 0, _drawablePointersBufferAddress = 0, _cullDataBufferAddress = 0;
	DrawableGpuData* _drawableStagingData = nullptr;
	DrawableCullData* _cullStagingData = nullptr;
	// recorded frame
	size_t _recordedDrawables = 0;
	bool _processingRecorded = false, _cullingRecorded = false;
	Frustum _frustum{};
	std::vector<DrawRange> _drawRanges;
	std::vector<uint64_t> _rangeInstances;   // worst-case instance count per draw range
	std::vector<uint64_t> _rangeCommands;
	uint64_t _rangeChunks = 0;
	CullResult _cull;
	size_t _cullCmdCapacity = 0, _cullInstCapacity = 0, _cullRangeCapacity = 0, _cullChunkCapacity = 0;
	uint64_t _cullRegionsAddress = 0, _cullWorkspaceAddress = 0;
	bool _collectFrameInfo = false;
	FrameInfo _inProgress, _completed;
	uint64_t _countsEpoch = 0;
	// optional pre-test of the culling pass (cadr_b200_compute_drawable_bounds): one box per flattened drawable,
	// recomputed by submit() whenever anything it depends on changed since it was last computed
	bool _useDrawableBounds = false;
	uint64_t _boundsAddress = 0, _boundsInputsEpoch = 0, _boundsComputedEpoch = ~uint64_t(0);
	size_t _boundsCapacity = 0;
	// device-resident drawable list: only ranges whose records changed since they were last copied are
	// written to the staging list and DMA'd; the reference re-copies the whole list every frame (Renderer.cpp:635-644)
	bool _incrementalList = true, _residentValid = false;
	std::vector<std::pair<size_t, size_t>> _dirtyRanges;   // [first, count) in drawables
	size_t _lastListUploadBytes = 0;
	static Renderer* _defaultRenderer;
	void freeDrawableBuffers() noexcept;
	void ensureCullBuffers();
	void recordStateSetRange(StateSet& ss, size_t firstDrawable);   // called by StateSet::recordToCommandBuffer
public:
	static constexpr size_t smallMemorySize = 64 << 10;
	static constexpr size_t mediumMemorySize = 2 << 20;
	static constexpr size_t largeMemorySize = 32 << 20;
	static constexpr uint32_t drawablePointersRecordSize = 4 * sizeof(uint64_t);
	static Renderer& get() { return *_defaultRenderer; }
	static void set(Renderer& r) noexcept { _defaultRenderer = &r; }

	/// device >= 0: CUDA device index (throws when no GPU is usable: there is no CPU fallback).
	/// device == addressSpaceOnly: host bookkeeping only, for tests of placement/handles/flattening on CPU boxes.
	static constexpr int addressSpaceOnly = -1;
	explicit Renderer(int device = 0, bool makeDefault = true);
	Renderer(cadr_ctx* ctx, void* stream, bool makeDefault = true);   ///< share a context / stream
	~Renderer();
	Renderer(const Renderer&) = delete;

	size_t beginFrame();
	void beginRecording();
	size_t prepareSceneRendering(StateSet& stateSetRoot);
	void recordDrawableProcessing(size_t numDrawables);
	void recordDrawableCulling(const Frustum& frustum);               ///< north-star extension; after recordSceneRendering
	void recordSceneRendering(StateSet& stateSetRoot);
	void endRecording();
	void submit();                                                    ///< stand-in for the app's vkQueueSubmit
	void waitIdle(uint64_t timeoutNs = 0);                            ///< fence wait; throws Timeout
	void endFrame();
	void executeCopyOperations();                                     ///< Renderer.cpp:946-999 (blocks, 1.5 s timeout)

	void notifyInstanceCountsChanged() noexcept { _countsEpoch++; }   ///< drawables / matrix-list sizes / LOD tables changed
	uint64_t countsEpoch() const { return _countsEpoch; }
	/// Anything the per-drawable bounds depend on changed: matrices, MatrixList of a drawable, model-space sphere,
	/// the set or order of drawables.
	void notifyBoundsInputsChanged() noexcept { _boundsInputsEpoch++; }
	/// Extension, off by default: let the culling pass drop long MatrixLists that lie outside the frustum before their
	/// matrices are read.  Results are identical; a frame after a scene change pays one extra pass over the matrices.
	void setDrawableBounds(bool on) { _useDrawableBounds = on; _boundsComputedEpoch = ~uint64_t(0); }
	bool hasDevice() const;
	/// Off: copy the whole flattened list every frame exactly like the reference.  On (default): copy what changed.
	void setIncrementalDrawableUpload(bool on) { _incrementalList = on; _residentValid = false; }
	size_t lastDrawableUploadBytes() const { return _lastListUploadBytes; }   ///< list + culling records DMA'd by the last submit()
	cadr_ctx* context() const { return _ctx; }
	void* stream() const { return _stream; }
	size_t frameNumber() const noexcept { return _frameNumber; }
	DataStorage& dataStorage() const { return *_dataStorage; }
	StagingManager& stagingManager() const { return *_stagingManager; }
	uint64_t drawableBufferAddress() const { return _drawableBufferAddress; }
	size_t drawableBufferSize() const { return _drawableCapacity * sizeof(DrawableGpuData); }
	DrawableGpuData* drawableStagingData() const { return _drawableStagingData; }
	DrawableCullData* cullStagingData() const { return _cullStagingData; }
	uint64_t drawIndirectBufferAddress() const { return _drawIndirectBufferAddress; }
	uint64_t drawablePointersBufferAddress() const { return _drawablePointersBufferAddress; }
	const std::vector<DrawRange>& drawRanges() const { return _drawRanges; }
	const CullResult& cullResult() const { return _cull; }
	bool collectFrameInfo() const { return _collectFrameInfo; }
	void setCollectFrameInfo(bool on);
	const FrameInfo& getFrameInfo();
	void readDevice(void* dst, uint64_t srcAddress, size_t bytes);    ///< blocking read-back (tests, tools)
};

}
