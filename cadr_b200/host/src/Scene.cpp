// Scene objects of the facade: Geometry, MatrixList, Drawable, StateSet (+ parent/child lists).
#include <CadR/CadR.h>
#include <algorithm>
#include <cstring>

namespace CadR {

// ---- Geometry ---------------------------------------------------------------------------------------
Geometry::Geometry(Renderer& r) : _vertices(r.dataStorage()), _indices(r.dataStorage()), _primitiveSets(r.dataStorage()) {}
Geometry::~Geometry()
{
	// drawables that still reference this geometry are detached (auto-unlink hooks in the reference)
	for(Drawable* d = _firstDrawable; d; ) {
		Drawable* next = d->_geometryNext;
		d->_geometry = nullptr; d->_geometryPrev = d->_geometryNext = nullptr;
		d = next;
	}
}

// ---- MatrixList -------------------------------------------------------------------------------------
MatrixList::MatrixList(Renderer& r) : _matrixList(r.dataStorage()) {}

void MatrixList::initHeader(void* p, size_t numMatrices)
{
	// {u32 numMatrices, u32 capacity, 56 zero bytes} (MatrixList.h:54)
	std::memset(p, 0, 64);
	uint32_t n = uint32_t(numMatrices);
	std::memcpy(p, &n, 4);
	std::memcpy(static_cast<uint8_t*>(p) + 4, &n, 4);
}

mat4* MatrixList::editNewContent(size_t numMatrices)
{
	StagingData sd = _matrixList.alloc(sizeof(mat4) * numMatrices + sizeof(mat4));
	mat4* m = sd.data<mat4>();
	initHeader(m, numMatrices);
	_matrixList.renderer().notifyBoundsInputsChanged();
	if(numMatrices != _numMatrices) {
		_numMatrices = numMatrices;
		_matrixList.renderer().notifyInstanceCountsChanged();
	}
	return m + 1;
}

// ---- Drawable ---------------------------------------------------------------------------------------
Drawable::Drawable(Geometry& geometry, uint32_t primitiveSetOffset, MatrixList& matrixList, StateSet& stateSet)
	: _matrixList(&matrixList), _drawableData(nullptr)
{
	linkToGeometry(geometry);
	stateSet.appendDrawableInternal(*this, DrawableGpuData(
		geometry.vertexDataAllocation().handle(), geometry.indexDataAllocation().handle(), matrixList.handle(),
		0, geometry.primitiveSetDataAllocation().handle(), primitiveSetOffset));
}

Drawable::Drawable(Geometry& geometry, uint32_t primitiveSetOffset, MatrixList& matrixList, DataAllocation& drawableData, StateSet& stateSet)
	: _matrixList(&matrixList), _drawableData(&drawableData)
{
	linkToGeometry(geometry);
	stateSet.appendDrawableInternal(*this, DrawableGpuData(
		geometry.vertexDataAllocation().handle(), geometry.indexDataAllocation().handle(), matrixList.handle(),
		drawableData.handle(), geometry.primitiveSetDataAllocation().handle(), primitiveSetOffset));
}

void Drawable::linkToGeometry(Geometry& geometry) noexcept
{
	_geometry = &geometry;
	_geometryPrev = nullptr;
	_geometryNext = geometry._firstDrawable;
	if(_geometryNext) _geometryNext->_geometryPrev = this;
	geometry._firstDrawable = this;
}

void Drawable::unlinkFromGeometry() noexcept
{
	if(_geometry) {
		if(_geometryPrev) _geometryPrev->_geometryNext = _geometryNext;
		else _geometry->_firstDrawable = _geometryNext;
		if(_geometryNext) _geometryNext->_geometryPrev = _geometryPrev;
		_geometry = nullptr;
		_geometryPrev = _geometryNext = nullptr;
	}
}

// this object replaces `other` in the list of other's geometry (move construction / assignment)
void Drawable::takeGeometryLinkOf(Drawable& other) noexcept
{
	_geometry = other._geometry; _geometryPrev = other._geometryPrev; _geometryNext = other._geometryNext;
	if(_geometry) {
		if(_geometryPrev) _geometryPrev->_geometryNext = this;
		else _geometry->_firstDrawable = this;
		if(_geometryNext) _geometryNext->_geometryPrev = this;
	}
	other._geometry = nullptr;
	other._geometryPrev = other._geometryNext = nullptr;
}

Drawable::~Drawable() noexcept
{
	if(_indexIntoStateSet != ~0u) _stateSet->removeDrawableInternal(*this);
	unlinkFromGeometry();
}

void Drawable::destroy() noexcept
{
	if(_indexIntoStateSet != ~0u) {
		_stateSet->removeDrawableInternal(*this);
		unlinkFromGeometry();
		_indexIntoStateSet = ~0u;
	}
}

Drawable::Drawable(Drawable&& o) noexcept
	: _stateSet(o._stateSet), _matrixList(o._matrixList), _drawableData(o._drawableData), _indexIntoStateSet(o._indexIntoStateSet)
{
	if(_indexIntoStateSet != ~0u) _stateSet->_drawablePtrList[_indexIntoStateSet] = this;
	takeGeometryLinkOf(o);
	o._indexIntoStateSet = ~0u;
}

Drawable& Drawable::operator=(Drawable&& rhs) noexcept
{
	if(this == &rhs) return *this;
	if(_indexIntoStateSet != ~0u) _stateSet->removeDrawableInternal(*this);
	unlinkFromGeometry();
	_stateSet = rhs._stateSet; _matrixList = rhs._matrixList; _drawableData = rhs._drawableData;
	_indexIntoStateSet = rhs._indexIntoStateSet;
	if(_indexIntoStateSet != ~0u) _stateSet->_drawablePtrList[_indexIntoStateSet] = this;
	takeGeometryLinkOf(rhs);
	rhs._indexIntoStateSet = ~0u;
	return *this;
}

void Drawable::create(Geometry& geometry, uint32_t primitiveSetOffset, MatrixList& matrixList, DataAllocation* drawableData, StateSet& stateSet)
{
	unlinkFromGeometry();
	linkToGeometry(geometry);
	_matrixList = &matrixList;
	_drawableData = drawableData;
	// NOTE: like the reference (Drawable.cpp:128,147) create() stores drawableDataHandle = 0 even when drawableData
	// is given; only the constructors pass the handle on.  Kept for result parity.
	DrawableGpuData gpuData(geometry.vertexDataAllocation().handle(), geometry.indexDataAllocation().handle(),
	                        matrixList.handle(), 0, geometry.primitiveSetDataAllocation().handle(), primitiveSetOffset);
	if(_indexIntoStateSet != ~0u) {
		if(_stateSet == &stateSet) {
			_stateSet->_drawableDataList[_indexIntoStateSet] = gpuData;
			_stateSet->_drawableCullList[_indexIntoStateSet].lodPrimitiveSetOffset[0] = primitiveSetOffset;
			_stateSet->_modCount++;
			_stateSet->renderer().notifyInstanceCountsChanged();
			return;
		}
		_stateSet->removeDrawableInternal(*this);
	}
	stateSet.appendDrawableInternal(*this, gpuData);
}

Renderer& Drawable::renderer() const { return _stateSet->renderer(); }

void Drawable::setCullData(const BoundingSphere& bs, uint32_t lodCount, const uint32_t* lodPrimitiveSetOffsets, const float* lodThresholds)
{
	if(_indexIntoStateSet == ~0u) throw LogicError("CadR::Drawable::setCullData(): the drawable is not attached to a StateSet");
	if(lodCount < 1 || lodCount > 3) throw LogicError("CadR::Drawable::setCullData(): lodCount must be 1, 2 or 3");
	DrawableCullData& c = _stateSet->_drawableCullList[_indexIntoStateSet];
	c.sphere[0] = bs.center[0]; c.sphere[1] = bs.center[1]; c.sphere[2] = bs.center[2]; c.sphere[3] = bs.radius;
	c.lodCount = lodCount;
	for(uint32_t l = 0; l < lodCount; l++)
		c.lodPrimitiveSetOffset[l] = lodPrimitiveSetOffsets ? lodPrimitiveSetOffsets[l] : _stateSet->_drawableDataList[_indexIntoStateSet].primitiveSetOffset;
	for(uint32_t l = 0; l + 1 < lodCount; l++)
		c.lodThreshold[l] = lodThresholds ? lodThresholds[l] : 0.f;
	_stateSet->_modCount++;
	_stateSet->renderer().notifyInstanceCountsChanged();
	_stateSet->renderer().notifyBoundsInputsChanged();
}

// ---- StateSet ---------------------------------------------------------------------------------------
StateSet::StateSet(Renderer& renderer) noexcept : _renderer(&renderer)
{
	childList._owner = this;
	parentList._owner = this;
}

static StateSetLink* linkStateSets(StateSet& parent, std::list<StateSetLink*>& childListOfParent, StateSet& child, std::list<StateSetLink*>& parentListOfChild)
{
	// ParentChildList.h:113-120 / :172-179: the relation is appended to BOTH lists
	StateSetLink* l = new StateSetLink{&parent, &child, {}, {}};
	l->inChildList = childListOfParent.insert(childListOfParent.end(), l);
	l->inParentList = parentListOfChild.insert(parentListOfChild.end(), l);
	return l;
}

template<> StateSetLinkList<true>::iterator StateSetLinkList<true>::append(StateSet& child)
{
	return linkStateSets(*_owner, _list, child, child.parentList._list)->inChildList;
}
template<> void StateSetLinkList<true>::remove(iterator it)
{
	StateSetLink* l = *it;
	l->child->parentList._list.erase(l->inParentList);
	_list.erase(l->inChildList);
	delete l;
}
template<> StateSetLinkList<false>::iterator StateSetLinkList<false>::append(StateSet& parent)
{
	return linkStateSets(parent, parent.childList._list, *_owner, _list)->inParentList;
}
template<> void StateSetLinkList<false>::remove(iterator it)
{
	StateSetLink* l = *it;
	l->parent->childList._list.erase(l->inChildList);
	_list.erase(l->inParentList);
	delete l;
}

void StateSet::appendDrawableInternal(Drawable& d, const DrawableGpuData& gpuData)
{
	d._stateSet = this;
	d._indexIntoStateSet = uint32_t(_drawableDataList.size());
	_drawableDataList.emplace_back(gpuData);
	DrawableCullData c{};
	c.sphere[3] = std::numeric_limits<float>::infinity();   // never culled until bounds are given
	c.lodCount = 1;
	c.lodPrimitiveSetOffset[0] = gpuData.primitiveSetOffset;
	_drawableCullList.push_back(c);
	_drawablePtrList.emplace_back(&d);
	_modCount++;
	_renderer->notifyInstanceCountsChanged();
}

void StateSet::removeDrawableInternal(Drawable& d) noexcept
{
	// swap-remove (StateSet.cpp:29-47)
	const uint32_t i = d._indexIntoStateSet;
	const size_t last = _drawableDataList.size() - 1;
	if(i != last) {
		_drawableDataList[i] = _drawableDataList[last];
		_drawableCullList[i] = _drawableCullList[last];
		Drawable* moved = _drawablePtrList[last];
		_drawablePtrList[i] = moved;
		moved->_indexIntoStateSet = i;
	}
	_drawableDataList.pop_back();
	_drawableCullList.pop_back();
	_drawablePtrList.pop_back();
	_modCount++;
	_renderer->notifyInstanceCountsChanged();
}

void StateSet::appendDrawable(Drawable& d, const DrawableGpuData& gpuData)
{
	if(d._indexIntoStateSet != ~0u) d._stateSet->removeDrawableInternal(d);
	appendDrawableInternal(d, gpuData);
}

void StateSet::removeDrawable(Drawable& d)
{
	if(d._indexIntoStateSet == ~0u) return;
	d._stateSet->removeDrawableInternal(d);
	d._indexIntoStateSet = ~0u;
}

void StateSet::removeAllDrawables() noexcept
{
	for(Drawable* d : _drawablePtrList) d->_indexIntoStateSet = ~0u;
	_drawableDataList.clear();
	_drawableCullList.clear();
	_drawablePtrList.clear();
	_modCount++;
	_renderer->notifyInstanceCountsChanged();
}

size_t StateSet::prepareRecording()
{
	_skipRecording = !_forceRecording;
	for(auto& f : prepareCallList) f(*this);
	size_t numDrawables = _drawableDataList.size();
	for(StateSet& ss : childList) {
		numDrawables += ss.prepareRecording();
		_skipRecording = _skipRecording && ss._skipRecording;
	}
	_skipRecording = _skipRecording && (numDrawables == 0);
	return numDrawables;
}

void StateSet::recordToCommandBuffer(size_t& drawableCounter)
{
	if(_skipRecording) return;   // subgraphs without drawables are not visited (StateSet.cpp:202-203)
	for(auto& f : recordCallList) f(*this, drawableCounter);
	const size_t numDrawables = _drawableDataList.size();
	if(numDrawables > 0) {
		_renderer->recordStateSetRange(*this, drawableCounter);
		drawableCounter += numDrawables;
	}
	for(StateSet& child : childList)
		child.recordToCommandBuffer(drawableCounter);
}

}
