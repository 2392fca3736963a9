// Drives a small dynamic scene through the facade exactly as an application drives CadR (call order of
// examples/RenderingPerformance/main.cpp:1424-1572) and dumps, per frame, everything needed to check it against
// the oracle from Python (tests/test_host_cpu.py, tests/test_host_gpu.py):
//   * the device image reconstructed from the upload regions the facade issued (upload observer),
//   * the flattened drawable list, culling records, draw ranges,
//   * what the facade's own getters say the result must be (expected indirect / pointers),
//   * with a device: what the GPU produced (Tier R outputs and the compacted Tier X buffers).
// usage: facade_scene_test <cuda device | -1> <dump file> [frames]
#include "frame_dump.h"
#include <memory>
#include <random>

using namespace CadR;

struct Geo { std::unique_ptr<Geometry> g; std::vector<PrimitiveSet> ps; };

int main(int argc, char** argv)
{
	if(argc < 3) { fprintf(stderr, "usage: %s <device|-1> <dump> [frames] [bounds]\n", argv[0]); return 2; }
	const int device = atoi(argv[1]);
	const int frames = argc > 3 ? atoi(argv[3]) : 4;
	dump::Writer w;
	w.out = fopen(argv[2], "wb");
	if(!w.out) return 2;
	try {
		Renderer r(device);
		if(argc > 4 && std::string(argv[4]) == "bounds") r.setDrawableBounds(true);
		dump::Shadow shadow;
		shadow.attach(r);
		std::mt19937 rng(1234);
		auto rnd = [&](uint32_t n) { return uint32_t(rng() % n); };

		StateSet root(r), a(r), b(r), shared(r), emptySet(r);
		root.childList.append(a);
		root.childList.append(b);
		a.childList.append(shared);
		b.childList.append(shared);        // two parents: recorded once per parent (StateSet.cpp:266-267)
		b.childList.append(emptySet);      // no drawables: skipped

		std::vector<Geo> geos;
		std::vector<std::unique_ptr<MatrixList>> lists;
		std::vector<std::unique_ptr<DataAllocation>> drawableData;
		std::vector<std::unique_ptr<DataAllocation>> filler;     // pushes the handle count over the level boundaries
		std::vector<std::unique_ptr<Drawable>> drawables;
		std::map<Drawable*, size_t> geoOf;
		std::map<Drawable*, uint32_t> psOffsetOf;

		auto fillList = [&](MatrixList& ml, size_t count) {
			mat4* m = ml.editNewContent(count);
			for(size_t i = 0; i < count; i++) {
				m[i] = mat4::translate(float(rnd(400)) - 200.f, float(rnd(400)) - 200.f, float(rnd(400)) - 200.f);
				float s = 0.5f + float(rnd(100)) / 50.f;
				m[i].m[0] = s; m[i].m[5] = s; m[i].m[10] = s;
			}
		};
		auto addDrawable = [&](size_t g, size_t l, StateSet& ss, bool withData) {
			uint32_t psOff = rnd(uint32_t(geos[g].ps.size())) * 8;
			std::unique_ptr<Drawable> d;
			if(withData) {
				drawableData.push_back(std::make_unique<DataAllocation>(r.dataStorage()));
				uint8_t payload[64]; for(auto& x : payload) x = uint8_t(rnd(256));
				drawableData.back()->setData(payload, sizeof(payload));
				d = std::make_unique<Drawable>(*geos[g].g, psOff, *lists[l], *drawableData.back(), ss);
			}
			else d = std::make_unique<Drawable>(*geos[g].g, psOff, *lists[l], ss);
			BoundingSphere bs{{float(rnd(5)) - 2.f, float(rnd(5)) - 2.f, float(rnd(5)) - 2.f}, rnd(10) == 0 ? -1.f : float(rnd(20))};
			uint32_t lods = 1 + rnd(3);
			uint32_t offs[3]; for(auto& o : offs) o = rnd(uint32_t(geos[g].ps.size())) * 8;
			float thr[2] = {float(50 + rnd(100)), float(150 + rnd(150))};
			d->setCullData(bs, lods, offs, thr);
			geoOf[d.get()] = g; psOffsetOf[d.get()] = psOff;
			drawables.push_back(std::move(d));
		};

		for(int frame = 0; frame < frames; frame++) {
			r.beginFrame();
			// ---- scene edits -------------------------------------------------------------------------
			if(frame == 0) {
				for(int g = 0; g < 6; g++) {
					Geo geo; geo.g = std::make_unique<Geometry>(r);
					std::vector<uint8_t> v(12 * (8 + rnd(20))), idx(4 * (6 + rnd(60)));
					for(auto& x : v) x = uint8_t(rnd(256));
					for(auto& x : idx) x = uint8_t(rnd(256));
					geo.g->uploadVertexData(v.data(), v.size());
					geo.g->uploadIndexData(idx.data(), idx.size());
					geo.ps.resize(1 + rnd(4));
					for(auto& p : geo.ps) p = PrimitiveSet{rnd(1000), rnd(1000)};
					geo.g->uploadPrimitiveSetData(geo.ps.data(), geo.ps.size() * sizeof(PrimitiveSet));
					geos.push_back(std::move(geo));
				}
				const size_t counts[] = {1, 0, 5, 40, 1500, 1, 33, 2, 1024, 1025};
				for(size_t c : counts) { lists.push_back(std::make_unique<MatrixList>(r)); fillList(*lists.back(), c); }
				StateSet* sets[] = {&root, &a, &b, &shared};
				for(int d = 0; d < 60; d++) addDrawable(rnd(6), rnd(10), *sets[rnd(4)], rnd(3) == 0);
			}
			if(frame == 1) {
				// rewrite transforms (realloc-on-write: new address, same handle), resize two lists
				fillList(*lists[2], 5); fillList(*lists[3], 7); fillList(*lists[4], 2100); fillList(*lists[1], 3);
				// remove drawables (swap-remove inside their StateSet) and re-associate one
				drawables[5].reset(); drawables[17].reset();
				drawables[20]->create(*geos[1].g, 0, *lists[0], shared);
				geoOf[drawables[20].get()] = 1; psOffsetOf[drawables[20].get()] = 0;
				// 2100 more handles: the table grows from one level to two (2047 -> 2048)
				for(int i = 0; i < 2100; i++) { filler.push_back(std::make_unique<DataAllocation>(r.dataStorage())); uint64_t v = i; filler.back()->setData(v); }
			}
			if(frame == 2) {
				for(int d = 0; d < 25; d++) addDrawable(rnd(6), rnd(10), d % 2 ? a : shared, d % 4 == 0);
				fillList(*lists[0], 1); fillList(*lists[9], 40);
				geos[2].ps[0] = PrimitiveSet{777, 42};
				geos[2].g->uploadPrimitiveSetData(geos[2].ps.data(), geos[2].ps.size() * sizeof(PrimitiveSet));
			}
			if(frame >= 3) {
				fillList(*lists[size_t(frame) % lists.size()], 10 + size_t(frame));
				if(frame == 3 && frames > 4) {
					// cross the second boundary (4194303 -> 4194304 handles): three table levels
					const size_t want = 4194400 - r.dataStorage().handleTable().highestHandle();
					for(size_t i = 0; i < want; i++) filler.push_back(std::make_unique<DataAllocation>(r.dataStorage(), DataAllocation::noHandle)), filler.back()->createHandle(r.dataStorage());
					addDrawable(0, 0, b, true);    // uses handles above 4194304
					lists.push_back(std::make_unique<MatrixList>(r)); fillList(*lists.back(), 3);
					addDrawable(1, lists.size() - 1, b, false);
				}
			}
			r.executeCopyOperations();

			// ---- record the frame --------------------------------------------------------------------
			r.beginRecording();
			const size_t n = r.prepareSceneRendering(root);
			r.recordDrawableProcessing(n);
			r.recordSceneRendering(root);
			Frustum f{};
			const float big = (frame % 2) ? 1e9f : 180.f;
			const float pl[6][4] = {{1, 0, 0, big}, {-1, 0, 0, big}, {0, 1, 0, big}, {0, -1, 0, big}, {0, 0, 1, big}, {0, 0, -1, big}};
			memcpy(f.planes, pl, sizeof(pl));
			f.eye[0] = 10.f; f.eye[1] = -20.f; f.eye[2] = 30.f;
			r.recordDrawableCulling(f);
			r.endRecording();
			r.executeCopyOperations();
			if(r.hasDevice()) { r.submit(); r.waitIdle(uint64_t(3e9)); }
			r.endFrame();

			// ---- dump ---------------------------------------------------------------------------------
			// what the facade itself says the processing result must be
			w.frame(r, shadow, n, f, frame, [&](Drawable& d, const DrawableGpuData& rec, uint32_t ind[4], uint64_t ptr[4]) {
				const Geo& geo = geos[geoOf[&d]];
				const PrimitiveSet& ps = geo.ps[psOffsetOf[&d] / 8];
				ind[0] = ps.indexCount; ind[1] = uint32_t(d.matrixList().numMatrices()); ind[2] = ps.startIndex; ind[3] = 0;
				ptr[0] = geo.g->vertexDataAllocation().deviceAddress(); ptr[1] = geo.g->indexDataAllocation().deviceAddress();
				ptr[2] = d.matrixList().allocation().deviceAddress();
				// Drawable::create() drops the drawable-data handle (reference behaviour, Drawable.cpp:128,147)
				ptr[3] = (d.drawableData() && rec.drawableDataHandle) ? d.drawableData()->deviceAddress() : 0;
			});
		}
		drawables.clear();
	}
	catch(Error& e) { fprintf(stderr, "CadR::Error: %s\n", e.what()); fclose(w.out); return 1; }
	fclose(w.out);
	return 0;
}
