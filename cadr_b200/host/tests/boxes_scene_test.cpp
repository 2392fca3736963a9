// The scenes of the reference's own system test, examples/RenderingPerformance (Tests.cpp: generateBakedBoxesScene
// :683-824, generateBoxesScene :456-618, updateDrawablesForShowHideScene :621-680, createTestScene :826-960), built
// with the facade the way that application builds them with CadR, at a test-sized grid, and dumped per frame for the
// oracle (tests/test_host_cpu.py on an address-space-only renderer, tests/test_host_gpu.py on the device):
//
//   baked               BakedBoxesScene: ONE geometry holding every box, one identity MatrixList, one drawable
//   instanced           InstancedBoxesScene: one box geometry, ONE MatrixList with a matrix per box, one drawable
//   independent         IndependentBoxesScene: a geometry, a one-matrix MatrixList and a drawable per box, one material
//   materials           IndependentBoxes1000MaterialsScene: the same with M materials handed out round-robin as
//                       per-drawable data (M = boxes / 8 here, 1000 in the reference)
//   material-per-box    IndependentBoxes1000000MaterialsScene: one material per box
//   shared-geometry     generateBoxesScene(instanced = false, singleGeometry = true): every drawable uses the ONE box
//                       geometry with its own one-matrix MatrixList (also the shape of BASELINE configs[1])
//   showhide            IndependentBoxesShowHideScene: no drawables at creation; EVERY frame all drawables are destroyed
//                       and every other box gets a new one, even boxes on even frames, odd ones on odd frames
//   showhide-shared     the singleGeometry branch of the same update (:667-671)
//   showhide-instanced  the instanced branch of the same update (:636-657): the one MatrixList is rewritten every frame
//                       with the matrices of every other column of boxes
//
// Every box gets its bounding sphere (Drawable::setCullData), so the culling extension runs on the same frames.
// usage: boxes_scene_test <cuda device | -1> <dump file | -> <scene> [boxes per side = 6] [frames = 3]
// With "-" instead of a dump file nothing is dumped and the host time of every phase of the frame is printed instead
// (the reference application's "scene update / construct" and "cpu time" columns, main.cpp:1634-1791).
#include "frame_dump.h"
#include <chrono>
#include <cmath>
#include <deque>
#include <string>

using namespace CadR;

struct MaterialData {            // examples/RenderingPerformance: 64 bytes of Phong parameters per material
	float ambient[3]; uint32_t padding1;
	float diffuse[3]; float alpha;
	float specular[3]; float shininess;
	float emission[3]; uint32_t padding2;
};
static_assert(sizeof(MaterialData) == 64, "MaterialData is 64 bytes");

// the 12 triangles of a box over corners numbered x + 2y + 4z
static const uint32_t kBoxIndices[36] = {
	0, 2, 1, 1, 2, 3,   0, 1, 4, 4, 1, 5,   0, 4, 2, 2, 4, 6,
	4, 5, 6, 6, 5, 7,   2, 6, 3, 3, 6, 7,   1, 3, 5, 5, 3, 7,
};

struct Grid {
	uint32_t nx, ny, nz;
	float spacing, boxSize;
	size_t count() const { return size_t(nx) * ny * nz; }
	// centre of box (i, j, k); the grid is centred on the origin
	void centre(uint32_t i, uint32_t j, uint32_t k, float c[3]) const {
		c[0] = (float(i) - float(nx - 1) / 2.f) * spacing;
		c[1] = (float(j) - float(ny - 1) / 2.f) * spacing;
		c[2] = (float(k) - float(nz - 1) / 2.f) * spacing;
	}
	template<typename F> void forEach(F&& f, uint32_t firstI = 0, uint32_t stepI = 1) const {   // x fastest, as in the reference
		for(uint32_t k = 0; k < nz; k++)
			for(uint32_t j = 0; j < ny; j++)
				for(uint32_t i = firstI; i < nx; i += stepI) f(i, j, k);
	}
};

static void writeCorners(float* p, const float c[3], float half)
{
	for(int v = 0; v < 8; v++) {
		p[3 * v + 0] = c[0] + ((v & 1) ? half : -half);
		p[3 * v + 1] = c[1] + ((v & 2) ? half : -half);
		p[3 * v + 2] = c[2] + ((v & 4) ? half : -half);
	}
}

struct Scene {
	Renderer& r;
	StateSet& root;
	Grid grid;
	std::deque<Geometry> geometries;
	std::vector<MatrixList> lists;
	std::vector<DataAllocation> materials;
	std::vector<Drawable> drawables;
	std::map<const Geometry*, PrimitiveSet> primitiveSetOf;   // what was uploaded for each geometry
	std::map<const Drawable*, const Geometry*> geometryOf;
	bool track = true;                 // off in timing mode: the maps above are test bookkeeping, not application work

	Scene(Renderer& r_, StateSet& root_, Grid g) : r(r_), root(root_), grid(g) {}

	Geometry& addBoxGeometry() {
		Geometry& g = geometries.emplace_back(r);
		const float origin[3] = {0.f, 0.f, 0.f};
		writeCorners(g.createVertexStagingData(8 * 12).data<float>(), origin, grid.boxSize / 2.f);
		memcpy(g.createIndexStagingData(sizeof(kBoxIndices)).data<uint32_t>(), kBoxIndices, sizeof(kBoxIndices));
		PrimitiveSet ps{36, 0};
		*g.createPrimitiveSetStagingData(sizeof(PrimitiveSet)).data<PrimitiveSet>() = ps;
		if(track) primitiveSetOf[&g] = ps;
		return g;
	}
	DataAllocation& addMaterial(size_t index, size_t of) {
		DataAllocation& a = materials.emplace_back(r);
		MaterialData& m = *a.editNewContent<MaterialData>(1);
		memset(&m, 0, sizeof(m));
		const float angle = float(index) / float(of), third = 1.04719755f;
		m.diffuse[0] = std::cos(angle - third); m.diffuse[1] = std::cos(angle); m.diffuse[2] = std::cos(angle + third);
		m.alpha = 1.f;
		return a;
	}
	void addDrawable(Geometry& g, MatrixList& ml, DataAllocation& material, float radius) {
		Drawable& d = drawables.emplace_back(g, 0, ml, material, root);
		d.setCullData(BoundingSphere{{0.f, 0.f, 0.f}, radius});
		if(track) geometryOf[&d] = &g;
	}
	float boxRadius() const { return std::sqrt(3.f) * grid.boxSize / 2.f; }

	// ---- creation ---------------------------------------------------------------------------------------------
	void createBaked() {
		const size_t n = grid.count();
		Geometry& g = geometries.emplace_back(r);
		float* pos = g.createVertexStagingData(n * 8 * 12).data<float>();
		uint32_t* idx = g.createIndexStagingData(n * 36 * 4).data<uint32_t>();
		size_t b = 0;
		grid.forEach([&](uint32_t i, uint32_t j, uint32_t k) {
			float c[3]; grid.centre(i, j, k, c);
			writeCorners(pos + b * 24, c, grid.boxSize / 2.f);
			for(int t = 0; t < 36; t++) idx[b * 36 + t] = uint32_t(b * 8) + kBoxIndices[t];
			b++;
		});
		PrimitiveSet ps{uint32_t(n * 36), 0};
		*g.createPrimitiveSetStagingData(sizeof(PrimitiveSet)).data<PrimitiveSet>() = ps;
		primitiveSetOf[&g] = ps;
		lists.reserve(1); materials.reserve(1); drawables.reserve(1);
		*lists.emplace_back(r).editNewContent(1) = mat4::identity();
		float far[3]; grid.centre(grid.nx - 1, grid.ny - 1, grid.nz - 1, far);
		const float h = grid.boxSize / 2.f;
		addDrawable(g, lists[0], addMaterial(0, 1), std::sqrt((far[0] + h) * (far[0] + h) + (far[1] + h) * (far[1] + h) + (far[2] + h) * (far[2] + h)));
	}
	void fillInstancedList(uint32_t firstI, uint32_t stepI) {
		// the reference sizes the rewritten list as ((nx - firstI) / 2) * ny * nz (Tests.cpp:643) but writes one matrix for
		// every i = firstI, firstI + 2, ... < nx, which is one more per row when nx - firstI is odd; here the list has
		// exactly the matrices that are written
		const size_t perRow = (grid.nx - firstI + stepI - 1) / stepI;
		mat4* m = lists[0].editNewContent(perRow * grid.ny * grid.nz);
		grid.forEach([&](uint32_t i, uint32_t j, uint32_t k) { float c[3]; grid.centre(i, j, k, c); *m++ = mat4::translate(c[0], c[1], c[2]); }, firstI, stepI);
	}
	void createInstanced() {
		addBoxGeometry();
		lists.reserve(1); materials.reserve(1); drawables.reserve(1);
		lists.emplace_back(r);
		fillInstancedList(0, 1);
		addDrawable(geometries.front(), lists[0], addMaterial(0, 1), boxRadius());
	}
	void createIndependent(size_t numMaterials, bool withDrawables, bool singleGeometry = false) {
		const size_t n = grid.count();
		for(size_t b = 0; b < (singleGeometry ? 1 : n); b++) addBoxGeometry();
		materials.reserve(numMaterials);
		for(size_t m = 0; m < numMaterials; m++) addMaterial(m, numMaterials);
		lists.reserve(n);
		grid.forEach([&](uint32_t i, uint32_t j, uint32_t k) {
			float c[3]; grid.centre(i, j, k, c);
			*lists.emplace_back(r).editNewContent(1) = mat4::translate(c[0], c[1], c[2]);
		});
		drawables.reserve(n);
		if(withDrawables)
			for(size_t b = 0; b < n; b++) addDrawable(geometries[singleGeometry ? 0 : b], lists[b], materials[b % numMaterials], boxRadius());
	}

	// ---- per-frame update of the show/hide scenes ---------------------------------------------------------------
	void showHideIndependent(size_t frameNumber) {
		geometryOf.clear();
		drawables.clear();                       // every Drawable is swap-removed from the StateSet
		// Tests.cpp:673-675: for(i = frame & 1, c = lists - i; i < c; i += 2) - the bound shrinks with the start index,
		// so an odd frame also leaves out the last box; the k-th new drawable takes geometry k, not geometry i
		size_t geometryIndex = 0;
		for(size_t i = frameNumber & 1, c = lists.size() - i; i < c; i += 2)
			addDrawable(geometries[geometries.size() == 1 ? 0 : geometryIndex++], lists[i], materials.back(), boxRadius());
	}
	void showHideInstanced(size_t frameNumber) { fillInstancedList(uint32_t(frameNumber & 1), 2); }
};

int main(int argc, char** argv)
{
	if(argc < 4) { fprintf(stderr, "usage: %s <device|-1> <dump> <scene> [boxes per side] [frames]\n", argv[0]); return 2; }
	const int device = atoi(argv[1]);
	const std::string kind = argv[3];
	const uint32_t side = argc > 4 ? uint32_t(atoi(argv[4])) : 6u;
	const int frames = argc > 5 ? atoi(argv[5]) : 3;
	const bool timingOnly = std::string(argv[2]) == "-";
	dump::Writer w;
	w.out = fopen(timingOnly ? "/dev/null" : argv[2], "wb");
	if(!w.out) return 2;
	auto now = [] { return std::chrono::steady_clock::now(); };
	auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
	try {
		Renderer r(device);
		dump::Shadow shadow;
		if(!timingOnly) shadow.attach(r);
		StateSet root(r);
		// the reference fits 100 boxes per side into 0.9 of the window's shorter edge; box size = half the spacing
		const float maxSize = 1080.f * 0.9f;
		Scene s(r, root, Grid{side, side > 1 ? side - 1 : 1, side > 2 ? side - 2 : 1, maxSize / float(side), maxSize / float(2 * side)});
		s.track = !timingOnly;
		if(timingOnly && r.hasDevice()) r.setCollectFrameInfo(true);   // per-kernel CUDA events of the library (FrameInfo)

		for(int frame = 0; frame < frames; frame++) {
			const auto t0 = now();
			r.beginFrame();
			if(frame == 0) {
				if(kind == "baked") s.createBaked();
				else if(kind == "instanced" || kind == "showhide-instanced") s.createInstanced();
				else if(kind == "independent") s.createIndependent(1, true);
				else if(kind == "materials") s.createIndependent(std::max<size_t>(s.grid.count() / 8, 2), true);
				else if(kind == "material-per-box") s.createIndependent(s.grid.count(), true);
				else if(kind == "shared-geometry") s.createIndependent(1, true, true);
				else if(kind == "showhide") s.createIndependent(1, false);
				else if(kind == "showhide-shared") s.createIndependent(1, false, true);
				else { fprintf(stderr, "unknown scene %s\n", kind.c_str()); return 2; }
			}
			// updateTestScene runs every frame, the first one included (main.cpp frame loop)
			if(kind == "showhide" || kind == "showhide-shared") s.showHideIndependent(r.frameNumber());
			if(kind == "showhide-instanced") s.showHideInstanced(r.frameNumber());
			const auto t1 = now();
			r.executeCopyOperations();
			const auto t2 = now();

			r.beginRecording();
			const size_t n = r.prepareSceneRendering(root);
			r.recordDrawableProcessing(n);
			r.recordSceneRendering(root);
			// an axis-aligned slab that cuts the grid in x on even frames (boxes straddling the cut stay), everything on odd ones
			Frustum f{};
			const float big = 1e9f, cut = (frame % 2) ? big : maxSize * 0.2f;
			const float pl[6][4] = {{1, 0, 0, cut}, {-1, 0, 0, big}, {0, 1, 0, big}, {0, -1, 0, big}, {0, 0, 1, big}, {0, 0, -1, big}};
			memcpy(f.planes, pl, sizeof(pl));
			f.eye[0] = 0.f; f.eye[1] = 0.f; f.eye[2] = -2.f * maxSize;
			r.recordDrawableCulling(f);
			r.endRecording();
			r.executeCopyOperations();
			if(r.hasDevice()) { r.submit(); r.waitIdle(uint64_t(3e9)); }
			r.endFrame();
			if(timingOnly) {
				const auto t3 = now();
				fprintf(stderr, "frame %d: %zu drawables, handle level %u | scene update %.2f ms, upload recording %.2f ms, frame recording %.2f ms, list upload %zu bytes\n",
				        frame, n, r.dataStorage().handleLevel(), ms(t0, t1), ms(t1, t2), ms(t2, t3), r.lastDrawableUploadBytes());
				if(r.hasDevice())     // the reference application's "gpu drawable processing" column (main.cpp:1717), here incl. culling
					fprintf(stderr, "         gpu drawable processing + culling %.4f ms\n", double(r.getFrameInfo().gpuEndExecution));
				continue;
			}

			w.frame(r, shadow, n, f, frame, [&](Drawable& d, const DrawableGpuData&, uint32_t ind[4], uint64_t ptr[4]) {
				const Geometry& g = *s.geometryOf.at(&d);
				const PrimitiveSet& ps = s.primitiveSetOf.at(&g);
				ind[0] = ps.indexCount; ind[1] = uint32_t(d.matrixList().numMatrices()); ind[2] = ps.startIndex; ind[3] = 0;
				ptr[0] = g.vertexDataAllocation().deviceAddress(); ptr[1] = g.indexDataAllocation().deviceAddress();
				ptr[2] = d.matrixList().allocation().deviceAddress();
				ptr[3] = d.drawableData() ? d.drawableData()->deviceAddress() : 0;
			});
		}
		// teardown newest first (an application that pops its drawables): unlinking from a shared geometry is O(1)
		const auto t0 = now();
		while(!s.drawables.empty()) s.drawables.pop_back();
		if(timingOnly) fprintf(stderr, "teardown of the drawables: %.2f ms\n", ms(t0, now()));
	}
	catch(Error& e) { fprintf(stderr, "CadR::Error: %s\n", e.what()); fclose(w.out); return 1; }
	catch(std::exception& e) { fprintf(stderr, "exception: %s\n", e.what()); fclose(w.out); return 1; }
	fclose(w.out);
	return 0;
}
