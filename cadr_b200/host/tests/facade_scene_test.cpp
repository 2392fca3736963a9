// Drives a small dynamic scene through the facade exactly as an application drives CadR (call order of
// examples/RenderingPerformance/main.cpp:1424-1572) and dumps, per frame, everything needed to check it against
// the oracle from Python (tests/test_host_cpu.py, tests/test_host_gpu.py):
//   * the device image reconstructed from the upload regions the facade issued (upload observer),
//   * the flattened drawable list, culling records, draw ranges,
//   * what the facade's own getters say the result must be (expected indirect / pointers),
//   * with a device: what the GPU produced (Tier R outputs and the compacted Tier X buffers).
// usage: facade_scene_test <cuda device | -1> <dump file> [frames]
#include <CadR/CadR.h>
#include "../../../include/cadr_b200.h"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <random>

using namespace CadR;

static FILE* g_out;
template<typename T> static void put(const T& v) { fwrite(&v, sizeof(T), 1, g_out); }
static void putBytes(const void* p, size_t n) { if(n) fwrite(p, 1, n, g_out); }

struct Shadow {
	std::map<uint64_t, std::vector<uint8_t>> seg;   // DataMemory base -> bytes
	void sync(DataStorage& ds) { for(DataMemory* m : ds.dataMemoryList()) if(m->size() && !seg.count(m->deviceAddress())) seg[m->deviceAddress()].assign(m->size(), 0); }
	void apply(const cadr_copy_region* r, size_t n) {
		for(size_t i = 0; i < n; i++) {
			auto it = seg.upper_bound(r[i].dstAddr);
			if(it == seg.begin()) { fprintf(stderr, "region outside every DataMemory\n"); exit(3); }
			--it;
			if(r[i].dstAddr + r[i].bytes > it->first + it->second.size()) { fprintf(stderr, "region overruns its DataMemory\n"); exit(3); }
			memcpy(it->second.data() + (r[i].dstAddr - it->first), reinterpret_cast<const void*>(r[i].srcOffset), r[i].bytes);
		}
	}
};

struct Geo { std::unique_ptr<Geometry> g; std::vector<PrimitiveSet> ps; };

int main(int argc, char** argv)
{
	if(argc < 3) { fprintf(stderr, "usage: %s <device|-1> <dump> [frames] [bounds]\n", argv[0]); return 2; }
	const int device = atoi(argv[1]);
	const int frames = argc > 3 ? atoi(argv[3]) : 4;
	g_out = fopen(argv[2], "wb");
	if(!g_out) return 2;
	try {
		Renderer r(device);
		if(argc > 4 && std::string(argv[4]) == "bounds") r.setDrawableBounds(true);
		Shadow shadow;
		r.dataStorage().uploadObserver = [&](const cadr_copy_region* regs, size_t n) { shadow.sync(r.dataStorage()); shadow.apply(regs, n); };
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
			shadow.sync(r.dataStorage());
			fwrite("CADRF002", 1, 8, g_out);
			put(uint32_t(frame)); put(uint32_t(device >= 0));
			put(uint32_t(shadow.seg.size()));
			for(auto& [base, bytes] : shadow.seg) { put(uint64_t(base)); put(uint64_t(bytes.size())); putBytes(bytes.data(), bytes.size()); }
			put(uint64_t(r.dataStorage().handleTableDeviceAddress())); put(uint32_t(r.dataStorage().handleLevel())); put(uint32_t(n));
			put(uint64_t(r.dataStorage().handleTable().highestHandle()));
			put(uint64_t(r.drawableBufferAddress()));
			putBytes(r.drawableStagingData(), n * 48);
			putBytes(r.cullStagingData(), n * 48);
			put(f);
			const auto& ranges = r.drawRanges();
			put(uint32_t(ranges.size()));
			for(size_t k = 0; k < ranges.size(); k++) {
				put(uint32_t(ranges[k].firstDrawable)); put(uint32_t(ranges[k].numDrawables));
				put(uint64_t(ranges[k].drawablePointersAddress - r.drawablePointersBufferAddress())); put(uint64_t(ranges[k].indirectOffset));
			}
			// what the facade itself says the processing result must be
			for(const DrawRange& dr : ranges)
				for(size_t i = 0; i < dr.numDrawables; i++) {
					Drawable& d = dr.stateSet->getDrawable(i);
					const Geo& geo = geos[geoOf[&d]];
					const PrimitiveSet& ps = geo.ps[psOffsetOf[&d] / 8];
					uint32_t ind[4] = {ps.indexCount, uint32_t(d.matrixList().numMatrices()), ps.startIndex, 0};
					// Drawable::create() drops the drawable-data handle (reference behaviour, Drawable.cpp:128,147)
					const DrawableGpuData& rec = dr.stateSet->drawableDataList()[i];
					uint64_t ptr[4] = {geo.g->vertexDataAllocation().deviceAddress(), geo.g->indexDataAllocation().deviceAddress(),
					                   d.matrixList().allocation().deviceAddress(),
					                   (d.drawableData() && rec.drawableDataHandle) ? d.drawableData()->deviceAddress() : 0};
					putBytes(ind, 16); putBytes(ptr, 32);
				}
			{
				const CullResult& c = r.cullResult();
				put(uint32_t(c.numRanges));
				for(auto& reg : c.regions) putBytes(reg.data(), 16);
			}
			if(r.hasDevice()) {
				std::vector<uint8_t> buf(n * 48);
				r.readDevice(buf.data(), r.drawIndirectBufferAddress(), n * 16); putBytes(buf.data(), n * 16);
				r.readDevice(buf.data(), r.drawablePointersBufferAddress(), n * 32); putBytes(buf.data(), n * 32);
				const CullResult& c = r.cullResult();
				uint64_t cmdCap = 0, instCap = 0;
				for(auto& reg : c.regions) { cmdCap += reg[1]; instCap += reg[3]; }
				put(cmdCap); put(instCap);
				std::vector<uint8_t> big2(std::max<size_t>({size_t(cmdCap) * 32, size_t(instCap) * 4, cadr_b200_cull_counters_bytes(c.numRanges), 16}));
				r.readDevice(big2.data(), c.counters, cadr_b200_cull_counters_bytes(c.numRanges)); putBytes(big2.data(), cadr_b200_cull_counters_bytes(c.numRanges));
				if(cmdCap) {
					r.readDevice(big2.data(), c.commands, cmdCap * 20); putBytes(big2.data(), cmdCap * 20);
					r.readDevice(big2.data(), c.pointers, cmdCap * 32); putBytes(big2.data(), cmdCap * 32);
					r.readDevice(big2.data(), c.tags, cmdCap * 8); putBytes(big2.data(), cmdCap * 8);
				}
				if(instCap) { r.readDevice(big2.data(), c.instances, instCap * 4); putBytes(big2.data(), instCap * 4); }
			}
			fprintf(stderr, "frame %d: %zu drawables, %zu ranges, handle level %u (%llu handles), %zu DataMemory, staging in use %zu / pooled %zu, list upload %zu bytes\n",
			        frame, n, ranges.size(), r.dataStorage().handleLevel(), (unsigned long long)r.dataStorage().handleTable().highestHandle(),
			        r.dataStorage().dataMemoryList().size(), r.stagingManager().numBlocksInUse(), r.stagingManager().numBlocksAvailable(),
			        r.lastDrawableUploadBytes());
		}
		drawables.clear();
	}
	catch(Error& e) { fprintf(stderr, "CadR::Error: %s\n", e.what()); fclose(g_out); return 1; }
	fclose(g_out);
	return 0;
}
