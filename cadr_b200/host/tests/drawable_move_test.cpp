// Drawable is movable (std::vector<CadR::Drawable> in the reference's applications): moving one must re-point the
// StateSet's back pointer and the Geometry's intrusive list of drawables at the new object; Geometry and StateSet may be
// destroyed before or after their drawables.  Exercised without reserve() so that the vector reallocates (move
// construction) and with erase() from the middle (move assignment).
#include <CadR/CadR.h>
#include <cstdio>
#include <memory>
#include <random>
#include <vector>

using namespace CadR;
#define REQUIRE(c) do { if(!(c)) { fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); return 1; } } while(0)

static bool backPointersConsistent(StateSet& ss, const std::vector<Drawable>& v)
{
	if(ss.getNumDrawables() != v.size()) return false;
	size_t found = 0;
	for(size_t i = 0; i < ss.getNumDrawables(); i++)
		for(const Drawable& d : v)
			if(&ss.getDrawable(i) == &d) { found++; break; }
	return found == v.size();
}

int main()
{
	Renderer r(Renderer::addressSpaceOnly);
	r.beginFrame();
	StateSet root(r);
	auto g1 = std::make_unique<Geometry>(r);
	auto g2 = std::make_unique<Geometry>(r);
	PrimitiveSet ps{36, 0};
	for(Geometry* g : {g1.get(), g2.get()}) {
		float v[24] = {};
		uint32_t idx[36] = {};
		g->uploadVertexData(v, sizeof(v)); g->uploadIndexData(idx, sizeof(idx)); g->uploadPrimitiveSetData(&ps, sizeof(ps));
	}
	std::vector<std::unique_ptr<MatrixList>> lists;
	for(int i = 0; i < 200; i++) { lists.push_back(std::make_unique<MatrixList>(r)); *lists.back()->editNewContent(1) = mat4::identity(); }

	// ---- move construction: the vector grows without reserve() ----------------------------------------
	std::vector<Drawable> drawables;
	for(int i = 0; i < 200; i++) drawables.emplace_back((i % 3) ? *g1 : *g2, 0, *lists[size_t(i)], root);
	REQUIRE(backPointersConsistent(root, drawables));
	for(size_t i = 0; i < drawables.size(); i++) REQUIRE(&drawables[i].matrixList() == lists[i].get() && drawables[i].isValid());

	// ---- move assignment: erase from the middle, in random places ----------------------------------------
	std::mt19937 rng(5);
	for(int k = 0; k < 120; k++) {
		const size_t at = rng() % drawables.size();
		MatrixList* next = at + 1 < drawables.size() ? &drawables[at + 1].matrixList() : nullptr;
		drawables.erase(drawables.begin() + long(at));
		if(next) REQUIRE(&drawables[at].matrixList() == next);
		REQUIRE(backPointersConsistent(root, drawables));
	}
	REQUIRE(drawables.size() == 80);
	// the records the StateSet holds still belong to the surviving drawables: one record per MatrixList handle left
	for(size_t i = 0; i < root.getNumDrawables(); i++)
		REQUIRE(root.drawableDataList()[i].matrixListHandle == root.getDrawable(i).matrixList().handle());

	// ---- a Geometry destroyed before its drawables: they are detached, destroying them later is harmless ----
	g2.reset();
	REQUIRE(backPointersConsistent(root, drawables));
	// moving a detached drawable and a linked one
	std::vector<Drawable> moved;
	for(Drawable& d : drawables) moved.push_back(std::move(d));
	drawables.clear();
	REQUIRE(backPointersConsistent(root, moved));
	// ---- drawables destroyed in random order, then the remaining Geometry --------------------------------------
	while(!moved.empty()) {
		const size_t at = rng() % moved.size();
		moved.erase(moved.begin() + long(at));
	}
	REQUIRE(root.getNumDrawables() == 0);
	g1.reset();
	// ---- a StateSet destroyed before its drawables (removeAllDrawables leaves them invalid, not dangling) ------
	{
		auto g3 = std::make_unique<Geometry>(r);
		float v[24] = {}; uint32_t idx[36] = {};
		g3->uploadVertexData(v, sizeof(v)); g3->uploadIndexData(idx, sizeof(idx)); g3->uploadPrimitiveSetData(&ps, sizeof(ps));
		std::vector<Drawable> tmp;
		{
			StateSet shortLived(r);
			for(int i = 0; i < 10; i++) tmp.emplace_back(*g3, 0, *lists[size_t(i)], shortLived);
			REQUIRE(shortLived.getNumDrawables() == 10);
		}
		for(Drawable& d : tmp) REQUIRE(!d.isValid());
		tmp.clear();
	}
	r.endFrame();
	printf("drawable_move_test ok\n");
	return 0;
}
