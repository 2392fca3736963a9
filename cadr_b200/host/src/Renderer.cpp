// Renderer of the facade: frame API over the C ABI.  See CadR/Renderer.h for the mapping to the reference.
#include <CadR/CadR.h>
#include <thread>
#include "../../../include/cadr_b200.h"
#include <chrono>
#include <cstring>

namespace CadR {

Renderer* Renderer::_defaultRenderer = nullptr;

static double now()
{
	return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

Renderer::Renderer(int device, bool makeDefault)
{
	if(device == addressSpaceOnly) check(cadr_b200_create_address_space_only(&_ctx));
	else check(cadr_b200_create(device, &_ctx));
	_ownsContext = true;
	_stream = cadr_b200_stream(_ctx);
	_stagingManager = std::make_unique<StagingManager>(_ctx);
	_dataStorage = std::make_unique<DataStorage>(*this);
	_dataStorage->init(*_stagingManager);
	if(makeDefault) _defaultRenderer = this;
}

Renderer::Renderer(cadr_ctx* ctx, void* stream, bool makeDefault) : _ctx(ctx), _stream(stream ? stream : cadr_b200_stream(ctx))
{
	_stagingManager = std::make_unique<StagingManager>(_ctx);
	_dataStorage = std::make_unique<DataStorage>(*this);
	_dataStorage->init(*_stagingManager);
	if(makeDefault) _defaultRenderer = this;
}

bool Renderer::hasDevice() const { return cadr_b200_device(_ctx) >= 0; }

Renderer::~Renderer()
{
	if(hasDevice()) cadr_b200_sync(_ctx, _stream, 0);
	_dataStorage.reset();       // handle table and buffers first: they return staging blocks
	_stagingManager.reset();
	freeDrawableBuffers();
	for(uint64_t a : {_cull.commands, _cull.pointers, _cull.tags, _cull.instances, _cull.counters, _cullRegionsAddress, _cullWorkspaceAddress, _boundsAddress})
		if(a) cadr_b200_arena_free(_ctx, a);
	if(_defaultRenderer == this) _defaultRenderer = nullptr;
	if(_ownsContext) cadr_b200_destroy(_ctx);
}

void Renderer::freeDrawableBuffers() noexcept
{
	for(uint64_t* a : {&_drawableBufferAddress, &_drawIndirectBufferAddress, &_drawablePointersBufferAddress, &_cullDataBufferAddress}) {
		if(*a) cadr_b200_arena_free(_ctx, *a);
		*a = 0;
	}
	if(_drawableStagingData) cadr_b200_host_free(_ctx, _drawableStagingData);
	if(_cullStagingData) cadr_b200_host_free(_ctx, _cullStagingData);
	_drawableStagingData = nullptr;
	_cullStagingData = nullptr;
	_drawableCapacity = 0;
}

size_t Renderer::beginFrame()
{
	_frameNumber++;
	// staging size hint = what the previous frame uploaded (Renderer.cpp:394-397)
	_lastFrameUploadBytes = _currentFrameUploadBytes;
	_currentFrameUploadBytes = 0;
	_dataStorage->setStagingDataSizeHint(_lastFrameUploadBytes);
	if(_collectFrameInfo) {
		_inProgress = FrameInfo{};
		_inProgress.frameNumber = _frameNumber;
		_inProgress.cpuBeginFrame = now();
	}
	return _frameNumber;
}

void Renderer::beginRecording()
{
	if(!_dirtyRanges.empty()) {
		// the previous recording was never submitted (or submit() threw before its copies were queued): its ranges were
		// noted as placed but never reached the device, and the staging they point into is about to be rewritten
		_dirtyRanges.clear();
		_residentValid = false;
	}
	_recordedDrawables = 0;
	_processingRecorded = _cullingRecorded = false;
	_drawRanges.clear();
	_rangeInstances.clear();
	_rangeCommands.clear();
	_rangeChunks = 0;
}

size_t Renderer::prepareSceneRendering(StateSet& stateSetRoot)
{
	if(_collectFrameInfo) _inProgress.cpuPrepareRecordingBegin = now();
	const size_t numDrawables = stateSetRoot.prepareRecording();
	if(_collectFrameInfo) _inProgress.cpuPrepareRecordingEnd = now();

	if(_drawableCapacity < numDrawables) {
		// 20 % head-room, at least 128 records; contents are per-frame and never preserved (Renderer.cpp:461-487)
		size_t n = size_t(float(numDrawables) * 1.2f);
		if(n < 128) n = 128;
		if(hasDevice()) cadr_b200_sync(_ctx, _stream, 0);
		freeDrawableBuffers();
		check(cadr_b200_arena_alloc(_ctx, n * sizeof(DrawableGpuData), &_drawableBufferAddress));
		void* p = nullptr;
		check(cadr_b200_host_alloc(_ctx, n * sizeof(DrawableGpuData), &p));
		_drawableStagingData = static_cast<DrawableGpuData*>(p);
		check(cadr_b200_arena_alloc(_ctx, n * sizeof(cadr_indirect_data), &_drawIndirectBufferAddress));
		check(cadr_b200_arena_alloc(_ctx, n * drawablePointersRecordSize, &_drawablePointersBufferAddress));
		check(cadr_b200_arena_alloc(_ctx, n * sizeof(DrawableCullData), &_cullDataBufferAddress));
		check(cadr_b200_host_alloc(_ctx, n * sizeof(DrawableCullData), &p));
		_cullStagingData = static_cast<DrawableCullData*>(p);
		_drawableCapacity = n;
		_residentValid = false;      // new buffers: nothing is resident
		_dirtyRanges.clear();        // ... and ranges noted against the old staging buffers mean nothing any more
	}
	return numDrawables;
}

void Renderer::recordDrawableProcessing(size_t numDrawables)
{
	if(numDrawables >= (size_t(1) << 30)) throw LogicError("Limit of 1Gi of Drawables reached.");   // Renderer.cpp:687
	_recordedDrawables = numDrawables;
	_processingRecorded = numDrawables != 0;    // nothing is dispatched for an empty scene (Renderer.cpp:600-620)
}

void StateSet::updateCullTotals()
{
	if(_totalsEpoch == _renderer->countsEpoch()) return;
	_instanceTotal = _commandTotal = _chunkTotal = 0;
	for(size_t i = 0; i < _drawablePtrList.size(); i++) {
		const uint64_t n = _drawablePtrList[i]->matrixList().numMatrices();
		const uint64_t lods = _drawableCullList[i].lodCount;
		_instanceTotal += n;
		if(n > CADR_CULL_SMALL_LIST_MAX) {    // becomes work items of the list kernel
			const uint64_t items = (n + CADR_CULL_WORK_ITEM_INSTANCES - 1) / CADR_CULL_WORK_ITEM_INSTANCES;
			_chunkTotal += items;
			_commandTotal += lods * items;
		}
		else {
			_commandTotal += n < lods ? n : lods;
		}
	}
	_totalsEpoch = _renderer->countsEpoch();
}

void Renderer::recordStateSetRange(StateSet& ss, size_t first)
{
	const size_t n = ss._drawableDataList.size();
	if(first + n > _drawableCapacity) throw LogicError("CadR::Renderer: more drawables recorded than prepareSceneRendering() counted");
	const uint32_t rangeIndex = uint32_t(_drawRanges.size());
	// was exactly this content copied to exactly this place before?  (k-th recording of the StateSet in the frame)
	if(ss._placementFrame != _frameNumber) { ss._placementFrame = _frameNumber; ss._placementCursor = 0; }
	const size_t k = ss._placementCursor++;
	if(ss._placements.size() <= k) ss._placements.resize(k + 1, StateSet::Placement{~size_t(0), ~0u, ~0ull});
	StateSet::Placement& pl = ss._placements[k];
	const bool resident = _incrementalList && _residentValid && pl.first == first && pl.range == rangeIndex && pl.modCount == ss._modCount;
	if(!resident) {
		// copy the StateSet's records into the staging list (StateSet.cpp:233-237)
		// (one core moves ~7 GB/s: at 10 M drawables the two copies take 145 ms, more than the PCIe transfer that
		// follows; large ranges are therefore cut into slices copied by several threads)
		auto copySlice = [&](size_t b, size_t e) {
			std::memcpy(&_drawableStagingData[first + b], ss._drawableDataList.data() + b, (e - b) * sizeof(DrawableGpuData));
			DrawableCullData* c = &_cullStagingData[first + b];
			std::memcpy(c, ss._drawableCullList.data() + b, (e - b) * sizeof(DrawableCullData));
			for(size_t i = 0; i < e - b; i++) c[i].stateSetIndex = rangeIndex;
		};
		constexpr size_t sliceMin = size_t(1) << 17;     // 128 Ki records = 12 MiB per slice at least
		const size_t threads = std::min<size_t>({n / sliceMin, size_t(std::max(1u, std::thread::hardware_concurrency())), size_t(16)});
		if(threads < 2) copySlice(0, n);
		else {
			std::vector<std::thread> pool;
			pool.reserve(threads - 1);
			const size_t per = (n + threads - 1) / threads;
			for(size_t t = 1; t < threads; t++) pool.emplace_back(copySlice, std::min(n, t * per), std::min(n, (t + 1) * per));
			copySlice(0, std::min(n, per));
			for(std::thread& t : pool) t.join();
		}
		if(!_dirtyRanges.empty() && _dirtyRanges.back().first + _dirtyRanges.back().second == first) _dirtyRanges.back().second += n;
		else _dirtyRanges.emplace_back(first, n);
		pl = StateSet::Placement{first, rangeIndex, ss._modCount};
	}
	_drawRanges.push_back(DrawRange{&ss, first, n, _drawablePointersBufferAddress + first * drawablePointersRecordSize,
	                                first * sizeof(cadr_indirect_data)});
	ss.updateCullTotals();
	_rangeInstances.push_back(ss._instanceTotal);
	_rangeCommands.push_back(ss._commandTotal);
	_rangeChunks += ss._chunkTotal;
}

void Renderer::recordSceneRendering(StateSet& stateSetRoot)
{
	if(_collectFrameInfo) _inProgress.cpuRecordStateSetsBegin = now();
	_drawRanges.clear();
	_rangeInstances.clear();
	_rangeCommands.clear();
	_rangeChunks = 0;
	size_t drawableCounter = 0;
	stateSetRoot.recordToCommandBuffer(drawableCounter);
	if(drawableCounter != _recordedDrawables && _processingRecorded)
		throw LogicError("CadR::Renderer::recordSceneRendering(): drawable count differs from recordDrawableProcessing()");
	if(_collectFrameInfo) _inProgress.cpuRecordStateSetsEnd = now();
}

void Renderer::recordDrawableCulling(const Frustum& frustum)
{
	_frustum = frustum;
	_cullingRecorded = true;
	// one output region per draw range, sized for the worst case (every instance visible, every LOD used)
	uint64_t cmds = 0, inst = 0;
	_cull.regions.resize(_drawRanges.size());
	for(size_t r = 0; r < _drawRanges.size(); r++) {
		_cull.regions[r] = {uint32_t(cmds), uint32_t(_rangeCommands[r]), uint32_t(inst), uint32_t(_rangeInstances[r])};
		cmds += _rangeCommands[r];
		inst += _rangeInstances[r];
	}
	if(cmds >= (1ull << 32) || inst >= (1ull << 32)) throw OutOfResources("CadR::Renderer: more than 4Gi instances in one culling pass");
	_cull.numRanges = uint32_t(_drawRanges.size());
}

void Renderer::endRecording() {}

void Renderer::ensureCullBuffers()
{
	uint64_t cmds = 0, inst = 0;
	for(auto& reg : _cull.regions) { cmds += reg[1]; inst += reg[3]; }
	auto grow = [&](uint64_t& addr, size_t& cap, size_t need, size_t elemBytes, std::initializer_list<std::pair<uint64_t*, size_t>> extra) {
		if(need <= cap && addr) return;
		size_t n = std::max<size_t>(size_t(double(need) * 1.2), 128);
		cadr_b200_sync(_ctx, _stream, 0);
		if(addr) cadr_b200_arena_free(_ctx, addr);
		addr = 0;
		check(cadr_b200_arena_alloc(_ctx, n * elemBytes, &addr));
		for(auto& e : extra) {
			if(*e.first) cadr_b200_arena_free(_ctx, *e.first);
			*e.first = 0;
			check(cadr_b200_arena_alloc(_ctx, n * e.second, e.first));
		}
		cap = n;
	};
	grow(_cull.commands, _cullCmdCapacity, cmds, sizeof(cadr_draw_indexed_indirect), {{&_cull.pointers, sizeof(cadr_drawable_pointers)}, {&_cull.tags, sizeof(cadr_command_tag)}});
	grow(_cull.instances, _cullInstCapacity, inst, sizeof(uint32_t), {});
	size_t rangeCap = _cullRangeCapacity;
	grow(_cullRegionsAddress, rangeCap, _drawRanges.size(), sizeof(cadr_stateset_region), {});
	if(rangeCap != _cullRangeCapacity || !_cull.counters) {
		if(_cull.counters) cadr_b200_arena_free(_ctx, _cull.counters);
		_cull.counters = 0;
		check(cadr_b200_arena_alloc(_ctx, cadr_b200_cull_counters_bytes(uint32_t(rangeCap)), &_cull.counters));
		_cullRangeCapacity = rangeCap;
	}
	grow(_cullWorkspaceAddress, _cullChunkCapacity, size_t(_rangeChunks), CADR_CULL_WORK_ITEM_BYTES, {});
}

void Renderer::submit()
{
	if(!_processingRecorded) return;
	if(!hasDevice()) throw DeviceError("CadR::Renderer::submit(): this renderer has no CUDA device; there is no CPU fallback");
	if(_collectFrameInfo) cadr_b200_set_profiling(_ctx, 1);
	// staging -> device copy of the flattened list (Renderer.cpp:635-644) — only the ranges that changed since they
	// were last copied; with incremental upload off, recordStateSetRange marked every range dirty
	_lastListUploadBytes = 0;
	_residentValid = false;              // until every copy below is queued (a throw leaves the list marked stale)
	for(auto& [first, count] : _dirtyRanges) {
		check(cadr_b200_memcpy_h2d(_ctx, _drawableBufferAddress + first * sizeof(DrawableGpuData), &_drawableStagingData[first],
		                           count * sizeof(DrawableGpuData), _stream));
		_lastListUploadBytes += count * sizeof(DrawableGpuData);
		if(_cullingRecorded) {
			check(cadr_b200_memcpy_h2d(_ctx, _cullDataBufferAddress + first * sizeof(DrawableCullData), &_cullStagingData[first],
			                           count * sizeof(DrawableCullData), _stream));
			_lastListUploadBytes += count * sizeof(DrawableCullData);
		}
	}
	_dirtyRanges.clear();
	_residentValid = _cullingRecorded;   // kept simple: residency is tracked across frames that upload both arrays
	if(!_cullingRecorded) {
		// the processing kernel (Renderer.cpp:669-692)
		check(cadr_b200_process_drawables(_ctx, _dataStorage->handleTableDeviceAddress(), _dataStorage->handleLevel(),
		                                  _drawableBufferAddress, _drawIndirectBufferAddress, _drawablePointersBufferAddress,
		                                  _recordedDrawables, _stream));
	}
	if(_cullingRecorded) {
		// culling recorded: the same DMA, then ONE pass that resolves handles, writes the indirect / pointers
		// records (identical to the processing kernel's) and culls
		ensureCullBuffers();
		check(cadr_b200_memcpy_h2d(_ctx, _cullRegionsAddress, _cull.regions.data(), _cull.regions.size() * sizeof(cadr_stateset_region), _stream));
		cadr_cull_params p{};
		p.handleTableRoot = _dataStorage->handleTableDeviceAddress();
		p.handleLevel = _dataStorage->handleLevel();
		p.numDrawables = uint32_t(_recordedDrawables);
		p.drawableList = _drawableBufferAddress;
		p.indirectData = _drawIndirectBufferAddress;
		p.drawablePointers = _drawablePointersBufferAddress;
		p.cullData = _cullDataBufferAddress;
		std::memcpy(p.planes, _frustum.planes, sizeof(p.planes));
		p.eye[0] = _frustum.eye[0]; p.eye[1] = _frustum.eye[1]; p.eye[2] = _frustum.eye[2];
		p.numStateSets = _cull.numRanges;
		p.stateSetRegions = _cullRegionsAddress;
		p.cmdOut = _cull.commands; p.ptrOut = _cull.pointers; p.tagOut = _cull.tags; p.instOut = _cull.instances;
		p.counters = _cull.counters;
		p.chunkWorkspace = _cullWorkspaceAddress;
		p.chunkCapacity = uint32_t(_rangeChunks);
		if(!_useDrawableBounds) {
			check(cadr_b200_process_and_cull(_ctx, &p, _stream));
			return;
		}
		// bounds are indexed by flattened drawable: every change of the list, of a MatrixList or of a sphere makes them stale
		const uint64_t epoch = _boundsInputsEpoch + _countsEpoch;
		if(_boundsCapacity < _recordedDrawables || !_boundsAddress) {
			if(_boundsAddress) check(cadr_b200_arena_free(_ctx, _boundsAddress));
			_boundsAddress = 0;
			_boundsCapacity = std::max<size_t>(size_t(double(_recordedDrawables) * 1.2), 128);
			check(cadr_b200_arena_alloc(_ctx, _boundsCapacity * sizeof(cadr_drawable_bound), &_boundsAddress));
			_boundsComputedEpoch = ~uint64_t(0);
		}
		if(_boundsComputedEpoch != epoch || _lastListUploadBytes != 0) {
			// two-call form: the bounds pass needs this frame's Tier R records (matrix list addresses, counts)
			check(cadr_b200_process_drawables(_ctx, p.handleTableRoot, p.handleLevel, p.drawableList, p.indirectData, p.drawablePointers,
			                                  _recordedDrawables, _stream));
			check(cadr_b200_compute_drawable_bounds(_ctx, &p, _boundsAddress, 0, uint32_t(_recordedDrawables), _stream));
			_boundsComputedEpoch = epoch;
			p.drawableBounds = _boundsAddress;
			check(cadr_b200_cull_compact(_ctx, &p, _stream));
		}
		else {
			p.drawableBounds = _boundsAddress;
			check(cadr_b200_process_and_cull(_ctx, &p, _stream));
		}
	}
}

void Renderer::waitIdle(uint64_t timeoutNs)
{
	if(hasDevice()) check(cadr_b200_sync(_ctx, _stream, timeoutNs));
}

void Renderer::endFrame()
{
	if(_collectFrameInfo) {
		_inProgress.cpuEndFrame = now();
		_completed = FrameInfo{};
	}
}

void Renderer::executeCopyOperations()
{
	auto [transferResources, numBytes] = _dataStorage->recordUploads(_stream);
	_currentFrameUploadBytes += numBytes;
	if(numBytes == 0) return;
	// the reference blocks on a fence with a 1.5 s timeout and throws CadR::Timeout (Renderer.cpp:971-993)
	if(hasDevice()) check(cadr_b200_sync(_ctx, _stream, uint64_t(1.5e9)));
	transferResources.release();
}

void Renderer::setCollectFrameInfo(bool on)
{
	_collectFrameInfo = on;
	if(hasDevice()) cadr_b200_set_profiling(_ctx, on ? 1 : 0);
}

const FrameInfo& Renderer::getFrameInfo()
{
	if(_collectFrameInfo && _completed.frameNumber != _inProgress.frameNumber && hasDevice()) {
		// drawable processing interval == ts[2] - ts[1] of the reference (main.cpp:1717)
		float ms[5] = {};               // KS_COUNT slots: process, cull small, cull list, scatter, patch
		check(cadr_b200_kernel_times(_ctx, ms, 5));
		_completed = _inProgress;
		_completed.gpuBeginExecution = 0.f;
		_completed.gpuAfterTransfersAndBeforeDrawableProcessing = 0.f;
		_completed.gpuAfterDrawableProcessingAndBeforeRendering = ms[0];
		_completed.gpuEndExecution = ms[0] + ms[1] + ms[2];
	}
	return _completed;
}

void Renderer::readDevice(void* dst, uint64_t srcAddress, size_t bytes)
{
	check(cadr_b200_memcpy_d2h(_ctx, dst, srcAddress, bytes, _stream));
	check(cadr_b200_sync(_ctx, _stream, 0));
}

}
