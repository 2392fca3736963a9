// facade_bench — the frame loop of the reference's system test (examples/RenderingPerformance/main.cpp:1374-1573:
// beginFrame .. executeCopyOperations .. beginRecording .. prepareSceneRendering .. recordDrawableProcessing ..
// recordSceneRendering .. endRecording .. executeCopyOperations .. submit .. endFrame) driven through CadR::Renderer of
// the facade on a real device, timed the way that application reports itself (main.cpp:1634-1791):
//   gpuDrawableProcessing   the device interval of the drawable-processing work (ts[2] - ts[1], main.cpp:1717), from the
//                           library's per-kernel events (Renderer::getFrameInfo)
//   cpu frame time          host time of one frame's API calls
// plus the end-to-end frame rate with the per-range counters read back to the host every frame.
// Scenes (BASELINE.json configs, built with the facade's own classes, nothing pre-baked):
//   c1   IndependentBoxesScene: side^3 boxes, a Geometry + one-matrix MatrixList + Drawable each      (Tests.cpp:456-618)
//   c2   one shared box Geometry, N drawables with a one-matrix MatrixList each, one StateSet          (configs[1])
//   c3   G geometries with 3 LOD PrimitiveSets x M-matrix MatrixLists over 64 StateSets                (configs[2])
//   c4   c3, and every frame 10 % of the MatrixLists are re-written through MatrixList::editNewContent (configs[3]):
//        realloc-on-write hands every rewritten list a NEW device range, the handle table follows (whole 16 KiB leaves are
//        re-staged and re-allocated, parents repointed), executeCopyOperations moves ~640 MB of staged matrices per frame.
//        The matrices written are the ones the lists already hold, so every frame must count exactly what the static
//        scene counts under the same camera - checked every frame: addresses move, results may not.
// usage: facade_bench <cuda device> <c1|c2|c3|c4> [frames = 200] [size: side | drawables | geometries] [matrices per list]
// Prints ONE JSON line.
#include <CadR/CadR.h>
#include "../../../include/cadr_b200.h"
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <memory>
#include <string>
#include <thread>
#include <vector>

using namespace CadR;

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// ---- counter-based PRNG (splitmix64), enough for a synthetic scene of the named shape ----------------------------------
static uint64_t mix(uint64_t x)
{
	x += 0x9E3779B97F4A7C15ull;
	x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
	x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
	return x ^ (x >> 31);
}
static float u01(uint64_t stream, uint64_t i) { return float(mix(i * 0xD1342543DE82EF95ull + stream * 0x9E3779B97F4A7C15ull) >> 40) * (1.f / 16777216.f); }
static float gauss(uint64_t stream, uint64_t i)
{
	const float a = std::max(u01(stream, i), 1e-7f), b = u01(stream + 1, i);
	return std::sqrt(-2.f * std::log(a)) * std::cos(6.2831853f * b);
}
// T * R * S, column-major
static mat4 trs(const float p[3], const float q[4], float s)
{
	const float x = q[0], y = q[1], z = q[2], w = q[3];
	mat4 r{};
	r.m[0] = (1 - 2 * (y * y + z * z)) * s; r.m[1] = (2 * (x * y + z * w)) * s; r.m[2] = (2 * (x * z - y * w)) * s;
	r.m[4] = (2 * (x * y - z * w)) * s; r.m[5] = (1 - 2 * (x * x + z * z)) * s; r.m[6] = (2 * (y * z + x * w)) * s;
	r.m[8] = (2 * (x * z + y * w)) * s; r.m[9] = (2 * (y * z - x * w)) * s; r.m[10] = (1 - 2 * (x * x + y * y)) * s;
	r.m[12] = p[0]; r.m[13] = p[1]; r.m[14] = p[2]; r.m[15] = 1.f;
	return r;
}

// ---- camera: eye on a circle around the origin, looking at it; LH, depth 0..1 (glm::lookAtLH, perspectiveLH_ZO) ----------
static Frustum orbitCamera(int frame, double radius, double farPlane, double fovyDeg = 60.0, double aspect = 16.0 / 9.0, double nearPlane = 0.5)
{
	const double a = (frame % 360) * 3.14159265358979323846 / 180.0;
	const double eye[3] = {radius * std::sin(a), 0.0, -radius * std::cos(a)};
	double f[3] = {-eye[0], -eye[1], -eye[2]};
	const double fl = std::sqrt(f[0] * f[0] + f[1] * f[1] + f[2] * f[2]);
	for(double& v : f) v /= fl;
	const double up[3] = {0, 1, 0};
	double s[3] = {up[1] * f[2] - up[2] * f[1], up[2] * f[0] - up[0] * f[2], up[0] * f[1] - up[1] * f[0]};
	const double sl = std::sqrt(s[0] * s[0] + s[1] * s[1] + s[2] * s[2]);
	for(double& v : s) v /= sl;
	const double u[3] = {f[1] * s[2] - f[2] * s[1], f[2] * s[0] - f[0] * s[2], f[0] * s[1] - f[1] * s[0]};
	double view[4][4] = {{s[0], s[1], s[2], -(s[0] * eye[0] + s[1] * eye[1] + s[2] * eye[2])},
	                     {u[0], u[1], u[2], -(u[0] * eye[0] + u[1] * eye[1] + u[2] * eye[2])},
	                     {f[0], f[1], f[2], -(f[0] * eye[0] + f[1] * eye[1] + f[2] * eye[2])},
	                     {0, 0, 0, 1}};
	const double t = std::tan(fovyDeg * 3.14159265358979323846 / 360.0);
	double proj[4][4] = {};
	proj[0][0] = 1 / (aspect * t); proj[1][1] = 1 / t; proj[2][2] = farPlane / (farPlane - nearPlane); proj[3][2] = 1;
	proj[2][3] = -(farPlane * nearPlane) / (farPlane - nearPlane);
	double pv[4][4] = {};
	for(int i = 0; i < 4; i++) for(int j = 0; j < 4; j++) for(int k = 0; k < 4; k++) pv[i][j] += proj[i][k] * view[k][j];
	// Gribb/Hartmann, depth 0..1: left, right, bottom, top, near, far
	double rows[6][4];
	for(int j = 0; j < 4; j++) {
		rows[0][j] = pv[3][j] + pv[0][j]; rows[1][j] = pv[3][j] - pv[0][j];
		rows[2][j] = pv[3][j] + pv[1][j]; rows[3][j] = pv[3][j] - pv[1][j];
		rows[4][j] = pv[2][j];            rows[5][j] = pv[3][j] - pv[2][j];
	}
	Frustum fr{};
	for(int k = 0; k < 6; k++) {
		const double n = std::sqrt(rows[k][0] * rows[k][0] + rows[k][1] * rows[k][1] + rows[k][2] * rows[k][2]);
		for(int j = 0; j < 4; j++) fr.planes[k][j] = float(rows[k][j] / n);
	}
	for(int j = 0; j < 3; j++) fr.eye[j] = float(eye[j]);
	return fr;
}

static const uint32_t kBoxIndices[72] = {
	0, 2, 1, 1, 2, 3,   0, 1, 4, 4, 1, 5,   0, 4, 2, 2, 4, 6,   4, 5, 6, 6, 5, 7,   2, 6, 3, 3, 6, 7,   1, 3, 5, 5, 3, 7,
	0, 2, 1, 1, 2, 3,   0, 1, 4, 4, 1, 5,   0, 4, 2, 2, 4, 6,   4, 5, 6, 6, 5, 7,                            // LOD 1: 24 indices from 36
	0, 2, 1, 1, 2, 3,   0, 1, 4, 4, 1, 5,                                                                    // LOD 2: 12 indices from 60
};

struct Scene {
	Renderer& r;
	StateSet root;
	std::deque<StateSet> sets;
	std::deque<Geometry> geometries;
	std::vector<MatrixList> lists;
	std::vector<Drawable> drawables;
	uint64_t instances = 0;
	std::vector<mat4> hostMatrices;      // c4: what every list holds, list after list
	size_t perList = 0;
	explicit Scene(Renderer& r_) : r(r_), root(r_) {}

	Geometry& addBox(bool lods) {
		Geometry& g = geometries.emplace_back(r);
		float* p = g.createVertexStagingData(8 * 12).data<float>();
		for(int v = 0; v < 8; v++) { p[3 * v] = (v & 1) ? 2.43f : -2.43f; p[3 * v + 1] = (v & 2) ? 2.43f : -2.43f; p[3 * v + 2] = (v & 4) ? 2.43f : -2.43f; }
		const size_t nIdx = lods ? 72 : 36;
		std::memcpy(g.createIndexStagingData(nIdx * 4).data<uint32_t>(), kBoxIndices, nIdx * 4);
		const PrimitiveSet ps[3] = {{36, 0}, {24, 36}, {12, 60}};          // SURVEY Appendix D cfg 3
		std::memcpy(g.createPrimitiveSetStagingData(lods ? sizeof(ps) : sizeof(PrimitiveSet)).data<PrimitiveSet>(), ps, lods ? sizeof(ps) : sizeof(PrimitiveSet));
		return g;
	}
};

static const float kBoxRadius = std::sqrt(3.f) * 2.43f;

int main(int argc, char** argv)
{
	if(argc < 3) { fprintf(stderr, "usage: facade_bench <cuda device> <c1|c2|c3> [frames] [size] [matrices per list]\n"); return 2; }
	const int device = atoi(argv[1]);
	const std::string which = argv[2];
	const int frames = argc > 3 ? atoi(argv[3]) : 200;
	try {
		Renderer r(device);
		Scene sc(r);
		const double tBuild0 = now();
		double radius = 1500.0, farPlane = 1500.0;
		std::string desc;
		if(which == "c1") {
			const uint32_t side = argc > 4 ? uint32_t(atoi(argv[4])) : 100;
			const size_t n = size_t(side) * side * side;
			sc.lists.reserve(n); sc.drawables.reserve(n);
			const float dist = 9.72f, origin = -dist * float(side - 1) / 2.f;
			size_t b = 0;
			for(uint32_t k = 0; k < side; k++) for(uint32_t j = 0; j < side; j++) for(uint32_t i = 0; i < side; i++, b++) {
				Geometry& g = sc.addBox(false);
				*sc.lists.emplace_back(r).editNewContent(1) = mat4::translate(origin + i * dist, origin + j * dist, origin + k * dist);
				sc.drawables.emplace_back(g, 0, sc.lists.back(), sc.root).setCullData(BoundingSphere{{0, 0, 0}, kBoxRadius});
				if((b & 0xffff) == 0xffff) r.executeCopyOperations();
			}
			sc.instances = n;
			radius = 600.0; farPlane = 1500.0;
			desc = "configs[0] IndependentBoxesScene " + std::to_string(side) + "^3 boxes (geometry + 1-matrix list + drawable each), one StateSet, orbiting perspective camera";
		}
		else if(which == "c2") {
			const size_t n = argc > 4 ? size_t(atoll(argv[4])) : 10000000;
			Geometry& g = sc.addBox(false);
			sc.lists.reserve(n); sc.drawables.reserve(n);
			for(size_t i = 0; i < n; i++) {
				const float p[3] = {(u01(0, i) - 0.5f) * 2000.f, (u01(1, i) - 0.5f) * 2000.f, (u01(2, i) - 0.5f) * 2000.f};
				const float s = 0.5f + 1.5f * u01(3, i);
				mat4 m = mat4::translate(p[0], p[1], p[2]);
				m.m[0] = m.m[5] = m.m[10] = s;
				*sc.lists.emplace_back(r).editNewContent(1) = m;
				sc.drawables.emplace_back(g, 0, sc.lists.back(), sc.root).setCullData(BoundingSphere{{0, 0, 0}, kBoxRadius});
				if((i & 0xfffff) == 0xfffff) r.executeCopyOperations();
			}
			sc.instances = n;
			desc = "configs[1] " + std::to_string(n) + " drawables x 1 matrix, one shared geometry, one StateSet, orbiting camera";
		}
		else if(which == "c3" || which == "c4") {
			const size_t G = argc > 4 ? size_t(atoll(argv[4])) : 100000;
			const size_t M = argc > 5 ? size_t(atoll(argv[5])) : 1000;
			const uint32_t S = 64;
			for(uint32_t s = 0; s < S; s++) { sc.sets.emplace_back(r); sc.root.childList.append(sc.sets.back()); }
			sc.lists.reserve(G); sc.drawables.reserve(G);
			sc.perList = M;
			if(which == "c4") sc.hostMatrices.resize(G * M);
			const uint32_t lodOff[3] = {0, 8, 16};
			const float lodThr[2] = {300.f, 900.f};
			for(size_t k = 0; k < G; k++) {
				Geometry& g = sc.addBox(true);
				mat4* m = sc.lists.emplace_back(r).editNewContent(M);
				const float c[3] = {(u01(0, k) - 0.5f) * 4000.f, (u01(1, k) - 0.5f) * 4000.f, (u01(2, k) - 0.5f) * 4000.f};
				for(size_t j = 0; j < M; j++) {
					const uint64_t gi = k * M + j;
					const float p[3] = {c[0] + 20.f * gauss(10, gi), c[1] + 20.f * gauss(12, gi), c[2] + 20.f * gauss(14, gi)};
					const float u1 = u01(20, gi), t2 = 6.2831853f * u01(21, gi), t3 = 6.2831853f * u01(22, gi);
					const float a = std::sqrt(1.f - u1), b = std::sqrt(u1);
					const float q[4] = {a * std::sin(t2), a * std::cos(t2), b * std::sin(t3), b * std::cos(t3)};
					m[j] = trs(p, q, 0.5f + 1.5f * u01(30, gi));
				}
				if(which == "c4") std::memcpy(&sc.hostMatrices[k * M], m, M * sizeof(mat4));
				sc.drawables.emplace_back(g, 0, sc.lists.back(), sc.sets[k % S]).setCullData(BoundingSphere{{0, 0, 0}, kBoxRadius}, 3, lodOff, lodThr);
				if((k & 0x1ff) == 0x1ff) r.executeCopyOperations();       // 512 lists = 32 MiB of staged matrices per transfer
			}
			sc.instances = uint64_t(G) * M;
			farPlane = 3000.0;
			desc = "configs[2] " + std::to_string(G) + " geometries x " + std::to_string(M) + "-matrix lists over 64 StateSets, 3 LODs, orbiting camera";
		}
		else { fprintf(stderr, "unknown scene %s\n", which.c_str()); return 2; }
		r.executeCopyOperations();
		const double buildSeconds = now() - tBuild0;

		// pinned read-back buffers for the per-range counters (what a renderer would feed to its indirect-count draws)
		cadr_ctx* ctx = r.context();
		void* pinned[2] = {nullptr, nullptr};
		size_t countersBytes = 0;

		auto frame = [&](int f, bool collect) {
			r.beginFrame();
			r.executeCopyOperations();
			r.beginRecording();
			const size_t n = r.prepareSceneRendering(sc.root);
			r.recordDrawableProcessing(n);
			r.recordSceneRendering(sc.root);
			r.recordDrawableCulling(orbitCamera(f, radius, farPlane));
			r.endRecording();
			r.executeCopyOperations();
			r.submit();
			r.endFrame();
			(void)collect;
			return n;
		};

		// warm-up: the first frames flatten and upload the whole list, grow the output buffers
		size_t numDrawables = 0;
		for(int f = 0; f < 5; f++) numDrawables = frame(f, false);
		r.waitIdle();
		countersBytes = cadr_b200_cull_counters_bytes(r.cullResult().numRanges);
		for(void*& p : pinned) if(cadr_b200_host_alloc(ctx, countersBytes, &p) != CADR_OK) throw std::runtime_error(cadr_b200_last_error());

		if(which == "c4") {
			// reference counters of the static scene under a fixed camera
			auto readCounters = [&]() { std::vector<uint8_t> c(countersBytes); r.readDevice(c.data(), r.cullResult().counters, countersBytes); return c; };
			frame(30, false); r.waitIdle();
			const std::vector<uint8_t> expect = readCounters();
			if(*reinterpret_cast<const uint32_t*>(expect.data())) throw std::runtime_error("culling pass reported a status");
			const size_t G = sc.lists.size(), M = sc.perList, rewrite = G / 10;
			std::vector<mat4*> dst(rewrite);
			std::vector<size_t> src(rewrite);
			const unsigned nThreads = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
			double fillMs = 0, allocMs = 0, copyMs = 0, gpuWaitMs = 0, totalMs = 0;
			uint64_t uploaded = 0;
			size_t level = 0;
			for(int f = 0; f < frames; f++) {
				const double t0 = now();
				r.beginFrame();
				for(size_t t = 0; t < rewrite; t++) {                    // SURVEY Appendix D cfg 4: lists {(f*10007 + t*7919) mod n}
					const size_t k = (size_t(f) * 10007 + t * 7919) % G;
					dst[t] = sc.lists[k].editNewContent(M);              // new device range, staging block handed out
					src[t] = k;
				}
				allocMs += (now() - t0) * 1e3;
				{   // the application's writes into staging, on all cores (the allocator calls above are single-threaded like the reference)
					std::vector<std::thread> pool;
					for(unsigned w = 1; w < nThreads; w++)
						pool.emplace_back([&, w]() { for(size_t t = w; t < rewrite; t += nThreads) std::memcpy(dst[t], &sc.hostMatrices[src[t] * M], M * sizeof(mat4)); });
					for(size_t t = 0; t < rewrite; t += nThreads) std::memcpy(dst[t], &sc.hostMatrices[src[t] * M], M * sizeof(mat4));
					for(auto& th : pool) th.join();
				}
				const double t1 = now();
				r.executeCopyOperations();                                // uploads (blocks like the reference)
				const double t2 = now();
				r.beginRecording();
				const size_t n = r.prepareSceneRendering(sc.root);
				r.recordDrawableProcessing(n);
				r.recordSceneRendering(sc.root);
				r.recordDrawableCulling(orbitCamera(30, radius, farPlane));
				r.endRecording();
				r.executeCopyOperations();
				r.submit();
				r.endFrame();
				r.waitIdle();
				const double t3 = now();
				const std::vector<uint8_t> got = readCounters();
				// status, near-band count and the per-range totals (not the queue cursors)
				if(std::memcmp(got.data(), expect.data(), 8) != 0 || std::memcmp(got.data() + 64, expect.data() + 64, countersBytes - 64) != 0)
					throw std::runtime_error("frame " + std::to_string(f) + ": the culled result changed although only addresses moved");
				fillMs += (t1 - t0) * 1e3; copyMs += (t2 - t1) * 1e3; gpuWaitMs += (t3 - t2) * 1e3; totalMs += (t3 - t0) * 1e3;
				uploaded = uint64_t(rewrite) * (M + 1) * sizeof(mat4);
				level = r.dataStorage().handleLevel();
			}
			size_t arenas = 0, arenaBytes = 0;
			for(const DataMemory* m : r.dataStorage().dataMemoryList()) { arenas++; arenaBytes += m->size(); }
			for(void* p : pinned) cadr_b200_host_free(ctx, p);
			printf("{\"bench\": \"facade_bench\", \"scene\": \"c4\", \"workload\": \"configs[3] %s; every frame %zu of the %zu MatrixLists re-written through MatrixList::editNewContent (realloc-on-write, handle table follows)\", "
			       "\"instances\": %llu, \"frames\": %d, \"results_identical_to_static_scene_every_frame\": true, "
			       "\"ms_per_frame\": %.3f, \"M_instances_per_s\": %.1f, \"host_fill_ms\": %.3f, \"of_which_realloc_on_write_calls_ms\": %.3f, \"executeCopyOperations_ms\": %.3f, \"record_submit_wait_ms\": %.3f, "
			       "\"staged_bytes_per_frame\": %llu, \"upload_GB_per_s\": %.1f, \"data_memories\": %zu, \"data_memory_bytes\": %zu, \"handle_level\": %zu, \"fill_threads\": %u}\n",
			       desc.c_str(), rewrite, G, (unsigned long long)sc.instances, frames, totalMs / frames, double(sc.instances) / (totalMs / frames * 1e-3) / 1e6,
			       fillMs / frames, allocMs / frames, copyMs / frames, gpuWaitMs / frames, (unsigned long long)uploaded, double(uploaded) / (copyMs / frames * 1e-3) / 1e9,
			       arenas, arenaBytes, level, nThreads);
			return 0;
		}

		// (1) device interval of the drawable processing + host time per frame, frame by frame (FrameInfo)
		r.setCollectFrameInfo(true);
		std::vector<double> gpuMs, cpuMs;
		uint64_t survivors = 0;
		for(int f = 0; f < std::min(frames, 50); f++) {
			const double t0 = now();
			frame(5 + f, true);
			cpuMs.push_back((now() - t0) * 1e3);
			r.waitIdle();
			const FrameInfo& fi = r.getFrameInfo();
			gpuMs.push_back(double(fi.gpuEndExecution - fi.gpuAfterTransfersAndBeforeDrawableProcessing));
			if(f == 0) {
				std::vector<uint8_t> c(countersBytes);
				r.readDevice(c.data(), r.cullResult().counters, countersBytes);
				const uint32_t status = *reinterpret_cast<const uint32_t*>(c.data());
				if(status) throw std::runtime_error("culling pass reported status " + std::to_string(status));
				for(uint32_t s = 0; s < r.cullResult().numRanges; s++)
					survivors += reinterpret_cast<const uint64_t*>(c.data() + 64)[s] >> 32;
			}
		}
		r.setCollectFrameInfo(false);
		std::sort(gpuMs.begin(), gpuMs.end()); std::sort(cpuMs.begin(), cpuMs.end());

		// (2) end to end, synchronised: every frame's counters are read back and waited for before the next frame starts
		r.waitIdle();
		double t0 = now();
		for(int f = 0; f < frames; f++) {
			frame(60 + f, false);
			if(cadr_b200_memcpy_d2h(ctx, pinned[0], r.cullResult().counters, countersBytes, r.stream()) != CADR_OK) throw std::runtime_error(cadr_b200_last_error());
			r.waitIdle();
		}
		const double syncMs = (now() - t0) * 1e3 / frames;

		// (3) end to end, queued: frames are recorded back to back (the host runs ahead like a renderer that records frame
		// k + 1 while frame k executes), counters of every frame copied to pinned memory, one wait at the end
		t0 = now();
		for(int f = 0; f < frames; f++) {
			frame(60 + f, false);
			if(cadr_b200_memcpy_d2h(ctx, pinned[f & 1], r.cullResult().counters, countersBytes, r.stream()) != CADR_OK) throw std::runtime_error(cadr_b200_last_error());
		}
		r.waitIdle();
		const double queuedMs = (now() - t0) * 1e3 / frames;
		for(void* p : pinned) cadr_b200_host_free(ctx, p);

		const double inst = double(sc.instances);
		printf("{\"bench\": \"facade_bench\", \"scene\": \"%s\", \"workload\": \"%s\", \"drawables\": %zu, \"instances\": %llu, \"frames\": %d, "
		       "\"build_seconds\": %.2f, \"survivor_fraction\": %.4f, "
		       "\"gpuDrawableProcessing_ms\": %.4f, \"gpuDrawableProcessing_M_instances_per_s\": %.1f, \"cpu_frame_ms\": %.4f, "
		       "\"e2e_synchronised\": {\"ms_per_frame\": %.4f, \"M_instances_per_s\": %.1f}, "
		       "\"e2e_queued\": {\"ms_per_frame\": %.4f, \"M_instances_per_s\": %.1f}, "
		       "\"list_upload_bytes_last_frame\": %zu, \"handle_level\": %u, "
		       "\"loop\": \"CadR::Renderer beginFrame/executeCopyOperations/beginRecording/prepareSceneRendering/recordDrawableProcessing/"
		       "recordSceneRendering/recordDrawableCulling/endRecording/executeCopyOperations/submit/endFrame (main.cpp:1374-1573); "
		       "gpuDrawableProcessing = median device interval of the processing+culling kernels (main.cpp:1717), cpu = median host time of the calls\"}\n",
		       which.c_str(), desc.c_str(), numDrawables, (unsigned long long)sc.instances, frames, buildSeconds, double(survivors) / inst,
		       gpuMs[gpuMs.size() / 2], inst / (gpuMs[gpuMs.size() / 2] * 1e-3) / 1e6, cpuMs[cpuMs.size() / 2],
		       syncMs, inst / (syncMs * 1e-3) / 1e6, queuedMs, inst / (queuedMs * 1e-3) / 1e6,
		       r.lastDrawableUploadBytes(), r.dataStorage().handleLevel());
		return 0;
	}
	catch(std::exception& e) {
		fprintf(stderr, "facade_bench FAILED: %s\n", e.what());
		return 1;
	}
}
