// Scene objects of the facade: PrimitiveSet, Geometry, MatrixList, Drawable, StateSet.
// Reference: src/CadR/{PrimitiveSet,Geometry,MatrixList,Drawable,StateSet,ParentChildList}.*
#pragma once
#include <CadR/DataAllocation.h>
#include <cmath>
#include <functional>
#include <limits>
#include <list>
#include <vector>

namespace CadR {

class Renderer;
class StateSet;
class Drawable;

struct PrimitiveSet {            // src/CadR/PrimitiveSet.h:12-15
	uint32_t indexCount;
	uint32_t startIndex;
};

/// Column-major 4x4 float matrix, layout-compatible with glm::mat4 (64 bytes).
struct mat4 {
	float m[16];
	static mat4 identity() { return mat4{{1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1}}; }
	static mat4 translate(float x, float y, float z) { mat4 r = identity(); r.m[12] = x; r.m[13] = y; r.m[14] = z; return r; }
};
static_assert(sizeof(mat4) == 64, "mat4 must be 64 bytes");

/// CadR::BoundingSphere (src/CadR/BoundingSphere.h:18-43): radius -inf means empty.
struct BoundingSphere {
	float center[3];
	float radius;
	static BoundingSphere empty() { return {{0.f, 0.f, 0.f}, -std::numeric_limits<float>::infinity()}; }
	static BoundingSphere infinite() { return {{0.f, 0.f, 0.f}, std::numeric_limits<float>::infinity()}; }
	bool isEmpty() const { return radius == -std::numeric_limits<float>::infinity(); }
};

/// DrawableGpuData (src/CadR/Drawable.h:32-43): the 48-byte record the processing kernel reads.
struct DrawableGpuData {
	uint64_t vertexDataHandle;
	uint64_t indexDataHandle;
	uint64_t matrixListHandle;
	uint64_t drawableDataHandle;
	uint64_t primitiveSetHandle;
	uint32_t primitiveSetOffset;
	uint32_t padding;
	DrawableGpuData() {}
	constexpr DrawableGpuData(uint64_t v, uint64_t i, uint64_t m, uint64_t d, uint64_t p, uint32_t off)
		: vertexDataHandle(v), indexDataHandle(i), matrixListHandle(m), drawableDataHandle(d), primitiveSetHandle(p),
		  primitiveSetOffset(off), padding(0) {}
};
static_assert(sizeof(DrawableGpuData) == 48, "DrawableGpuData is 48 bytes (Drawable.cpp:13-15)");

/// Per-drawable culling record of the north-star extension (include/cadr_b200.h: cadr_drawable_cull_data).
struct DrawableCullData {
	float    sphere[4];
	uint32_t lodCount;
	uint32_t lodPrimitiveSetOffset[3];
	float    lodThreshold[2];
	uint32_t stateSetIndex;     ///< filled while the scene is flattened
	uint32_t reserved;
};
static_assert(sizeof(DrawableCullData) == 48, "cull record is 48 bytes");

class Geometry {                  // src/CadR/Geometry.h
	friend class Drawable;
	DataAllocation _vertices, _indices, _primitiveSets;
	Drawable* _firstDrawable = nullptr;   ///< intrusive list of the drawables that use this geometry (O(1) link / unlink, like the reference's auto-unlink hooks)
public:
	explicit Geometry(Renderer& r);
	Geometry(const Geometry&) = delete;
	~Geometry();
	Renderer& renderer() const { return _vertices.renderer(); }
	DataStorage& dataStorage() const { return _vertices.dataStorage(); }
	size_t vertexDataSize() const { return _vertices.size(); }
	size_t indexDataSize() const { return _indices.size(); }
	size_t primitiveSetDataSize() const { return _primitiveSets.size(); }
	DataAllocation& vertexDataAllocation() { return _vertices; }
	const DataAllocation& vertexDataAllocation() const { return _vertices; }
	DataAllocation& indexDataAllocation() { return _indices; }
	const DataAllocation& indexDataAllocation() const { return _indices; }
	DataAllocation& primitiveSetDataAllocation() { return _primitiveSets; }
	const DataAllocation& primitiveSetDataAllocation() const { return _primitiveSets; }
	void uploadVertexData(const void* ptr, size_t numBytes) { _vertices.setData(ptr, numBytes); }
	void uploadIndexData(const void* ptr, size_t numBytes) { _indices.setData(ptr, numBytes); }
	void uploadPrimitiveSetData(const void* ptr, size_t numBytes) { _primitiveSets.setData(ptr, numBytes); }
	StagingData createVertexStagingData(size_t numBytes) { return _vertices.alloc(numBytes); }
	StagingData createIndexStagingData(size_t numBytes) { return _indices.alloc(numBytes); }
	StagingData createPrimitiveSetStagingData(size_t numBytes) { return _primitiveSets.alloc(numBytes); }
	void freeVertexData() { _vertices.free(); }
	void freeIndexData() { _indices.free(); }
	void freePrimitiveSetData() { _primitiveSets.free(); }
	void freeData() { freeVertexData(); freeIndexData(); freePrimitiveSetData(); }
};

class MatrixList {                // src/CadR/MatrixList.h: 64-B header {numMatrices, capacity, 0...} + N x mat4
	DataAllocation _matrixList;
	size_t _numMatrices = 0;
	static void initHeader(void* p, size_t numMatrices);
public:
	explicit MatrixList(Renderer& r);
	MatrixList(MatrixList&&) = default;
	template<typename M> void setMatrices(const M& matrix) { static_assert(sizeof(M) == 64, "mat4 expected"); setMatrices(&matrix, 1); }
	template<typename M> void setMatrices(const std::vector<M>& matrices) { static_assert(sizeof(M) == 64, "mat4 expected"); setMatrices(matrices.data(), matrices.size()); }
	template<typename M> void setMatrices(const M* matrices, size_t numMatrices) {
		static_assert(sizeof(M) == 64, "mat4 expected");
		std::memcpy(editNewContent(numMatrices), matrices, 64 * numMatrices);
	}
	mat4* editNewContent(size_t numMatrices);
	uint64_t handle() const { return _matrixList.handle(); }
	size_t numMatrices() const { return _numMatrices; }
	const DataAllocation& allocation() const { return _matrixList; }
};

class Drawable {                  // src/CadR/Drawable.{h,cpp}
	friend class StateSet;
	friend class Geometry;
	StateSet* _stateSet = nullptr;
	MatrixList* _matrixList = nullptr;
	DataAllocation* _drawableData = nullptr;
	Geometry* _geometry = nullptr;
	uint32_t _indexIntoStateSet = ~0u;
	Drawable* _geometryPrev = nullptr;     // neighbours in Geometry's list of drawables
	Drawable* _geometryNext = nullptr;
	void create(Geometry& geometry, uint32_t primitiveSetOffset, MatrixList& matrixList, DataAllocation* drawableData, StateSet& stateSet);
	void linkToGeometry(Geometry& geometry) noexcept;
	void unlinkFromGeometry() noexcept;
	void takeGeometryLinkOf(Drawable& other) noexcept;
public:
	Drawable() noexcept = default;
	Drawable(Geometry& geometry, uint32_t primitiveSetOffset, MatrixList& matrixList, StateSet& stateSet);
	Drawable(Geometry& geometry, uint32_t primitiveSetOffset, MatrixList& matrixList, DataAllocation& drawableData, StateSet& stateSet);
	Drawable(const Drawable&) = delete;
	Drawable(Drawable&& other) noexcept;
	~Drawable() noexcept;
	Drawable& operator=(const Drawable&) = delete;
	Drawable& operator=(Drawable&& rhs) noexcept;
	void create(Geometry& geometry, uint32_t primitiveSetOffset, MatrixList& matrixList, StateSet& stateSet) { create(geometry, primitiveSetOffset, matrixList, nullptr, stateSet); }
	void create(Geometry& geometry, uint32_t primitiveSetOffset, MatrixList& matrixList, DataAllocation& drawableData, StateSet& stateSet) { create(geometry, primitiveSetOffset, matrixList, &drawableData, stateSet); }
	void destroy() noexcept;
	bool isValid() const { return _indexIntoStateSet != ~0u; }
	Renderer& renderer() const;
	StateSet& stateSet() const { return *_stateSet; }
	MatrixList& matrixList() const { return *_matrixList; }
	DataAllocation* drawableData() const { return _drawableData; }
	/// North-star extension: model-space bounds and LOD table used by Renderer::recordDrawableCulling.
	/// Without it a drawable is never culled (infinite sphere) and has the single LOD of its PrimitiveSet.
	void setCullData(const BoundingSphere& bs, uint32_t lodCount = 1, const uint32_t* lodPrimitiveSetOffsets = nullptr,
	                 const float* lodThresholds = nullptr);
};

/// Parent/child lists of the StateSet DAG (src/CadR/ParentChildList.h): a StateSet may have several parents and
/// is then recorded once per parent (StateSet.cpp:266-267).  As in the reference, every append creates ONE relation
/// object that sits in the parent's child list and in the child's parent list; removing it from either side removes
/// exactly that relation from both (also when the same two StateSets are linked more than once), in O(1).
struct StateSetLink {
	StateSet* parent;
	StateSet* child;
	std::list<StateSetLink*>::iterator inChildList;    ///< position in parent->childList
	std::list<StateSetLink*>::iterator inParentList;   ///< position in child->parentList
};

template<bool IsChildList> class StateSetLinkList {
	friend class StateSet;
	template<bool> friend class StateSetLinkList;
	StateSet* _owner = nullptr;
	std::list<StateSetLink*> _list;
	static StateSet& other(const StateSetLink* l) { return IsChildList ? *l->child : *l->parent; }
public:
	using iterator = std::list<StateSetLink*>::iterator;
	struct deref_iterator {
		std::list<StateSetLink*>::const_iterator it;
		StateSet& operator*() const { return other(*it); }
		StateSet* operator->() const { return &other(*it); }
		deref_iterator& operator++() { ++it; return *this; }
		deref_iterator& operator--() { --it; return *this; }
		bool operator!=(const deref_iterator& o) const { return it != o.it; }
		bool operator==(const deref_iterator& o) const { return it == o.it; }
	};
	deref_iterator begin() const { return {_list.begin()}; }
	deref_iterator end() const { return {_list.end()}; }
	size_t size() const { return _list.size(); }
	bool empty() const { return _list.empty(); }
	StateSet& front() const { return other(_list.front()); }
	StateSet& back() const { return other(_list.back()); }
	iterator append(StateSet& other);
	void remove(iterator it);
	void remove(deref_iterator it) { remove(_list.erase(it.it, it.it)); }   // the iterator begin() hands out, as in the reference
	void clear() { while(!_list.empty()) remove(_list.begin()); }
	~StateSetLinkList() { clear(); }
};

class StateSet {                  // src/CadR/StateSet.{h,cpp} (Vulkan pipeline/descriptor state is out of scope)
	friend class Drawable;
	friend class Renderer;
	template<bool> friend class StateSetLinkList;
	Renderer* _renderer;
	bool _skipRecording = true;
	bool _forceRecording = false;
	std::vector<DrawableGpuData> _drawableDataList;
	std::vector<DrawableCullData> _drawableCullList;   // parallel to _drawableDataList
	std::vector<Drawable*> _drawablePtrList;
	// worst-case output sizes of the culling extension for this StateSet, recomputed when instance counts change
	uint64_t _totalsEpoch = ~0ull, _instanceTotal = 0, _commandTotal = 0, _chunkTotal = 0;
	void updateCullTotals();
	// device-resident drawable list: where this StateSet's records were placed the last time they were copied
	// (one entry per recording of the StateSet in a frame: a StateSet with several parents is recorded several times)
	uint64_t _modCount = 0;                  ///< bumped whenever _drawableDataList / _drawableCullList change
	struct Placement { size_t first; uint32_t range; uint64_t modCount; };
	std::vector<Placement> _placements;
	size_t _placementFrame = ~size_t(0), _placementCursor = 0;
	void appendDrawableInternal(Drawable& d, const DrawableGpuData& gpuData);
	void removeDrawableInternal(Drawable& d) noexcept;
public:
	/// Callbacks, as in the reference: prepareCallList runs in prepareRecording(); recordCallList runs when the
	/// StateSet is recorded and receives the index of its first drawable in the flattened list.
	std::vector<std::function<void(StateSet&)>> prepareCallList;
	std::vector<std::function<void(StateSet&, size_t firstDrawable)>> recordCallList;
	StateSetLinkList<true> childList;
	StateSetLinkList<false> parentList;

	explicit StateSet(Renderer& renderer) noexcept;
	~StateSet() noexcept { destroy(); }
	void destroy() noexcept { removeAllDrawables(); }
	StateSet(const StateSet&) = delete;
	StateSet& operator=(const StateSet&) = delete;

	Renderer& renderer() const { return *_renderer; }
	bool forceRecording() const { return _forceRecording; }
	void setForceRecording(bool value) { _forceRecording = value; }
	void requestRecording() { _skipRecording = false; }

	size_t prepareRecording();                                               // StateSet.cpp:180-195
	void recordToCommandBuffer(size_t& drawableCounter);                     // StateSet.cpp:198-268

	void appendDrawable(Drawable& d, const DrawableGpuData& gpuData);
	static void removeDrawable(Drawable& d);
	void removeAllDrawables() noexcept;
	Drawable& getDrawable(size_t index) const { return *_drawablePtrList[index]; }
	size_t getNumDrawables() const { return _drawablePtrList.size(); }
	const std::vector<DrawableGpuData>& drawableDataList() const { return _drawableDataList; }
};

}
