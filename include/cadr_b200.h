/* cadr_b200.h — C ABI of libcadr_b200.so, the B200 (sm_100a) backend for CADR's per-frame
 * drawable processing and data upload.
 *
 * The reference (Rendering-FIT/CADR) has no FFI/plugin seam of its own: CadR is a C++ class library
 * whose GPU work is recorded into Vulkan command buffers.  The two places where control crosses from
 * scene bookkeeping to GPU work are
 *     CadR::Renderer::recordDrawableProcessing   src/CadR/Renderer.cpp:598-720
 *     CadR::Renderer::executeCopyOperations      src/CadR/Renderer.cpp:946-999
 * and everything the GPU needs there is (a) the 32-byte push-constant block
 * {handleTableRoot, drawableListPtr, indirectDataPtr, drawablePointersBufferPtr}
 * (src/CadR/shaders/processDrawables.comp:68-74, Renderer.cpp:677-682) and (b) the copy-region
 * triples {srcOffset, dstOffset, size} (src/CadR/DataMemory.cpp:417-425).  This ABI is cut exactly
 * there.  Every entry point cites the reference interface it replaces.
 *
 * Conventions
 *   - plain C, no torch / CUDA types in signatures; a stream is an opaque `void*` holding a
 *     cudaStream_t (NULL = the context's own stream);
 *   - device addresses are raw 64-bit CUDA device pointers and play the role of VkDeviceAddress;
 *   - every call returns CADR_OK (0) or a negative error; the message is kept per thread and read with
 *     cadr_b200_last_error().  Error taxonomy mirrors src/CadR/Exceptions.h:13-40
 *     (LogicError / OutOfResources / Timeout) plus CADR_E_CUDA for driver/runtime failures;
 *   - a context is bound to one GPU; calls on one context must be externally serialised (the reference
 *     is single-threaded: no mutex/atomic anywhere in src/CadR); all device work is stream-ordered and
 *     asynchronous, cadr_b200_sync() is the fence wait;
 *   - there is NO CPU fallback: without a CUDA device every compute entry fails with CADR_E_NO_DEVICE.
 */
#ifndef CADR_B200_H
#define CADR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
# define CADR_API
#else
# define CADR_API __attribute__((visibility("default")))
#endif

#define CADR_B200_ABI_VERSION 6

/* ---- error codes (src/CadR/Exceptions.h:13-40) ------------------------------------------------- */
enum {
	CADR_OK                  =  0,
	CADR_E_LOGIC             = -1,  /* CadR::LogicError: bad argument / API misuse                 */
	CADR_E_OUT_OF_RESOURCES  = -2,  /* CadR::OutOfResources: device/host allocation failed          */
	CADR_E_TIMEOUT           = -3,  /* CadR::Timeout: fence wait expired (Renderer.cpp:989-993)      */
	CADR_E_CUDA              = -4,  /* vk::Error analogue: CUDA runtime/driver error                 */
	CADR_E_NO_DEVICE         = -5,  /* no CUDA device: compute entry called on an address-only ctx   */
	CADR_E_OVERFLOW          = -6   /* an output region given to cadr_b200_cull_compact was too small */
};

typedef struct cadr_ctx cadr_ctx;
typedef void* cadr_stream;  /* cudaStream_t */

/* ---- device layouts (little-endian, std430) ---------------------------------- */

/* DrawableGpuData, 48 B — src/CadR/Drawable.h:32-43, processDrawables.comp:17-27 */
typedef struct cadr_drawable_gpu_data {
	uint64_t vertexDataHandle;
	uint64_t indexDataHandle;
	uint64_t matrixListHandle;
	uint64_t drawableDataHandle;   /* 0 = none (looked up like any other: slot 0 is zero)          */
	uint64_t primitiveSetHandle;
	uint32_t primitiveSetOffset;   /* BYTES into the PrimitiveSet array                            */
	uint32_t padding;
} cadr_drawable_gpu_data;

/* PrimitiveSet, 8 B — src/CadR/PrimitiveSet.h:12-15, processDrawables.comp:29-33 */
typedef struct cadr_primitive_set { uint32_t count, first; } cadr_primitive_set;

/* IndirectData == VkDrawIndirectCommand, 16 B — processDrawables.comp:43-50, Renderer.cpp:467 */
typedef struct cadr_indirect_data {
	uint32_t vertexCount, instanceCount, firstVertex, baseInstance;
} cadr_indirect_data;

/* DrawablePointers, 32 B — processDrawables.comp:52-59, Renderer.h:201 */
typedef struct cadr_drawable_pointers {
	uint64_t vertexDataPtr, indexDataPtr, matrixListPtr, drawableDataPtr;
} cadr_drawable_pointers;

/* MatrixList block: 64-B header {u32 numMatrices, u32 capacity, 56 B zero} + N x mat4 (column-major
 * f32) — src/CadR/MatrixList.h:54-59, processDrawables.comp:35-41 */
#define CADR_MATRIX_LIST_HEADER_BYTES 64u
#define CADR_MATRIX_BYTES             64u

/* Handle table node: 2048 x u64 — src/CadR/HandleTable.h:33-35 */
#define CADR_HANDLES_PER_TABLE   2048u
#define CADR_HANDLE_LEVEL_SHIFT  11u
#define CADR_HANDLE_LEVEL_MASK   0x7ffu

/* ---- Tier X (north-star extension; NOT present in the reference: parity unpinned) --------------- */

/* VkDrawIndexedIndirectCommand, 20 B */
typedef struct cadr_draw_indexed_indirect {
	uint32_t indexCount, instanceCount, firstIndex;
	int32_t  vertexOffset;
	uint32_t firstInstance;     /* index of the first entry of this run in the instance-index buffer */
} cadr_draw_indexed_indirect;

/* Per-drawable culling record, 48 B, array parallel to the drawable list.
 * sphere = CadR::BoundingSphere {vec3 center, float radius} in model space
 * (src/CadR/BoundingSphere.h:18-21); radius < 0 means empty => never visible (:39-43). */
typedef struct cadr_drawable_cull_data {
	float    sphere[4];
	uint32_t lodCount;                   /* 1..3                                                    */
	uint32_t lodPrimitiveSetOffset[3];   /* bytes into the PrimitiveSet array, like primitiveSetOffset */
	float    lodThreshold[2];            /* ascending eye distances; lod = #(threshold <= dist)      */
	uint32_t stateSetIndex;              /* index into cadr_cull_params::stateSetRegions             */
	uint32_t reserved;
} cadr_drawable_cull_data;

/* Output region of one StateSet (element indices into cmdOut/ptrOut/tagOut resp. instOut). */
typedef struct cadr_stateset_region {
	uint32_t cmdBase, cmdCapacity, instBase, instCapacity;
} cadr_stateset_region;

/* {drawableIndex, lod} of an emitted command — used for canonical sorting and by the consumer. */
typedef struct cadr_command_tag { uint32_t drawableIndex, lod; } cadr_command_tag;

/* Head of the counters buffer; followed by numStateSets packed u64:
 *   low 32 bits  = number of commands emitted for the StateSet (usable directly as the count buffer
 *                  of vkCmdDrawIndexedIndirectCount),
 *   high 32 bits = number of instance indices emitted for the StateSet. */
typedef struct cadr_cull_header {
	uint32_t status;          /* bit 0: a region overflowed, bit 1: chunk workspace overflowed, bit 2: bad range index */
	uint32_t nearBandCount;   /* instances whose deciding plane test, or LOD threshold, is within 1e-5 */
	uint32_t chunkCount;      /* internal: work items queued for the list kernel                     */
	uint32_t chunkCursor;     /* internal: work items claimed                                        */
	uint32_t medCount;        /* internal: medium lists queued (same workspace, filled from its end)  */
	uint32_t medCursor;       /* internal: medium lists claimed                                      */
	uint32_t reserved[10];
} cadr_cull_header;           /* 64 B */
#define CADR_CULL_STATUS_REGION_OVERFLOW 1u
#define CADR_CULL_STATUS_CHUNK_OVERFLOW  2u
#define CADR_CULL_STATUS_BAD_RANGE_INDEX 4u     /* a culling record's stateSetIndex >= numStateSets: drawable skipped */
#define CADR_CULL_STATUS_EXCHANGE_TIMEOUT 8u    /* multi-GPU: a peer did not publish its frame within cadr_exchange_sync::timeoutMs */
#define CADR_CULL_WORK_ITEM_BYTES        128u   /* one self-contained descriptor per <= 1024 matrices */
#define CADR_CULL_SMALL_LIST_MAX         32u    /* lists up to this size are evaluated by one thread; longer ones become work items */
#define CADR_CULL_WORK_ITEM_INSTANCES    1024u
#define CADR_CULL_MEDIUM_LIST_MAX        64u    /* lists of 33..64 matrices are one work item each, consumed 32 at a time */
#define CADR_CULL_BOUNDS_MIN_LIST         4u     /* the optional bounds pre-test applies to lists of at least this many matrices */

#define CADR_MAX_PEERS 8

typedef struct cadr_cull_params {
	/* Tier R inputs, same meaning as the push constants (processDrawables.comp:68-74) */
	uint64_t handleTableRoot;
	uint32_t handleLevel;        /* 1..3                                                            */
	uint32_t numDrawables;
	uint64_t drawableList;       /* cadr_drawable_gpu_data[numDrawables]                            */
	/* Tier R outputs of this frame, consumed here (numMatrices, matrixListPtr, pointers to forward) */
	uint64_t indirectData;       /* cadr_indirect_data[numDrawables]                                */
	uint64_t drawablePointers;   /* cadr_drawable_pointers[numDrawables]                            */
	/* Tier X inputs */
	uint64_t cullData;           /* cadr_drawable_cull_data[numDrawables]                           */
	float    planes[6][4];       /* world-space (nx,ny,nz,d), inward-facing, unit normals           */
	float    eye[4];             /* camera position xyz, w unused                                   */
	uint32_t numStateSets;
	uint32_t reserved0;
	uint64_t stateSetRegions;    /* cadr_stateset_region[numStateSets]                              */
	/* outputs */
	uint64_t cmdOut;             /* cadr_draw_indexed_indirect[], 20-B stride                       */
	uint64_t ptrOut;             /* cadr_drawable_pointers[], parallel to cmdOut                    */
	uint64_t tagOut;             /* cadr_command_tag[], parallel to cmdOut                          */
	uint64_t instOut;            /* uint32_t instance indices                                       */
	uint64_t counters;           /* cadr_cull_header + uint64_t[numStateSets]; zeroed by the call    */
	/* scratch */
	uint64_t chunkWorkspace;     /* CADR_CULL_WORK_ITEM_BYTES per work item of the list kernel, 16-B aligned */
	uint32_t chunkCapacity;      /* >= sum over drawables with > CADR_CULL_SMALL_LIST_MAX matrices of ceil(numMatrices/1024) */
	uint32_t reserved1;
	/* Multi-GPU, fused exchange (no reference counterpart).  With exchangeWorld >= 2 the kernels store
	 * every emitted command / pointers / tag record straight into the gathered arrays of ALL ranks over NVLink peer
	 * mappings (slot exchangeRank * exchangeCmdCapacity + index) while the cull is still running; cmdOut / ptrOut /
	 * tagOut are then ignored.  exchangeCmd[r] etc. are rank r's gathered arrays as mapped in THIS process
	 * (cadr_b200_ipc_import; entry [exchangeRank] is the local buffer).  exchangeWorld 0 or 1: no exchange. */
	uint32_t exchangeWorld;
	uint32_t exchangeRank;
	uint32_t exchangeCmdCapacity;
	uint32_t reserved2;
	uint64_t exchangeCmd[CADR_MAX_PEERS];
	uint64_t exchangePtr[CADR_MAX_PEERS];
	uint64_t exchangeTag[CADR_MAX_PEERS];
	/* Optional pre-test (no reference counterpart): cadr_drawable_bound[numDrawables] written by
	 * cadr_b200_compute_drawable_bounds, or 0.  A drawable with at least CADR_CULL_BOUNDS_MIN_LIST matrices whose bound
	 * lies outside one frustum plane by more than a rounding-safe margin is dropped before any of its matrices is read;
	 * the result of the frame is identical with and without the table. */
	uint64_t drawableBounds;
	/* Consumer side only (cadr_b200_consume_check_culled; the cull ignores it): added to every forwarded pointer before it
	 * is dereferenced.  A renderer GPU that walks ANOTHER rank's commands reads that rank's geometry and matrix lists
	 * through a peer mapping (cadr_b200_ipc_import); the records hold addresses of the owning GPU, and
	 * addressDelta = (address of the owner's arena as mapped here) - (its address on the owner) translates them
	 * (two's complement: the mapping may lie below).  0 for local results. */
	uint64_t addressDelta;
} cadr_cull_params;

/* World-space axis-aligned box that encloses the bounding spheres of ALL instances of a drawable (the drawable's
 * model-space sphere under every matrix of its MatrixList), slightly inflated; valid < 0: no bound (never used to
 * skip).  Must be recomputed when the drawable's matrices, MatrixList or model-space sphere change. */
typedef struct cadr_drawable_bound { float center[3]; float valid; float halfExtent[3]; float reserved; } cadr_drawable_bound;  /* 32 B */

/* Publishing a rank's per-range counters to its peers and signalling "frame complete" (stream-ordered after the
 * cull kernels), and waiting until every peer has done the same.  Flags are monotonically increasing frame
 * sequence numbers, one u64 per rank in each rank's flag array. */
typedef struct cadr_exchange_sync {
	uint32_t world, rank;
	uint64_t frameSeq;                        /* > 0, strictly increasing per call pair                      */
	uint64_t localCounters;                   /* this rank's counters buffer (header + packed counts)         */
	uint32_t countersBytes;
	uint32_t timeoutMs;                       /* 0: the waits spin until every peer has published (default); else the budget of a wait
	                                           * in ms - when it is spent the wait ends and CADR_CULL_STATUS_EXCHANGE_TIMEOUT is set in
	                                           * the status word of localCounters (the device-side form of the reference's bounded
	                                           * fence wait -> CadR::Timeout, Renderer.cpp:982-993)                                   */
	uint64_t peerCounters[CADR_MAX_PEERS];    /* rank r's gathered counters [world][countersBytes] as mapped here */
	uint64_t peerFlags[CADR_MAX_PEERS];       /* rank r's flag array [world] u64 as mapped here                  */
} cadr_exchange_sync;

/* Renderer-side gather of the survivors' instance-index runs of every rank (optional second stage of the exchange:
 * with it a renderer reads instance indices locally and only the matrices through peer mappings).  For every
 * (rank r, range s) the run [instBase, instBase + count) of rank r's instance-index buffer - count = high half of the
 * counter rank r published - is copied to gatheredInst + (r * instCapacity + instBase) * 4. */
typedef struct cadr_exchange_pull {
	uint32_t world, rank;
	uint32_t numRanges, countersBytes;
	uint64_t gatheredCounters;                /* local gathered counters [world][countersBytes]                 */
	uint64_t gatheredInst;                    /* local uint32_t [world][instCapacity], 16-byte aligned          */
	uint64_t instCapacity;                    /* elements per rank, a multiple of 4                             */
	uint32_t includeLocal, reserved;          /* 0: skip this rank's own runs (they are local already)          */
	uint64_t regions[CADR_MAX_PEERS];         /* device copy of rank r's cadr_stateset_region[numRanges]        */
	uint64_t peerInst[CADR_MAX_PEERS];        /* rank r's instance-index buffer as mapped here                  */
} cadr_exchange_pull;

/* ---- upload ------------------------------------------------------------------ */

/* One copy region == one vk::BufferCopy recorded by DataMemory::recordUploads
 * (src/CadR/DataMemory.cpp:417-425): src is an offset into the staging block, dst a device address. */
typedef struct cadr_copy_region {
	uint64_t dstAddr;
	uint64_t srcOffset;
	uint64_t bytes;
} cadr_copy_region;

/* One handle-table update == HandleTable::set(handle, addr) (src/CadR/HandleTable.cpp:348-378). */
typedef struct cadr_handle_patch { uint64_t handle, addr; } cadr_handle_patch;

/* ---- context ------------------------------------------------------------------------------------ */

CADR_API int         cadr_b200_abi_version(void);
CADR_API const char* cadr_b200_last_error(void);

/* Replaces VulkanLibrary::load + VulkanInstance::chooseDevice + VulkanDevice::create + Renderer::init
 * (src/CadR/Renderer.cpp:97-313): binds the context to CUDA device `device`, creates its stream and
 * scratch buffers.  Fails with CADR_E_NO_DEVICE when no usable GPU exists. */
CADR_API int  cadr_b200_create(int device, cadr_ctx** out);

/* Address-space-only context for host-logic tests on machines without a GPU: arena_alloc hands out
 * addresses from a fake range, host_alloc uses plain memory, and every entry that would touch a device
 * returns CADR_E_NO_DEVICE.  It computes nothing. */
CADR_API int  cadr_b200_create_address_space_only(cadr_ctx** out);

CADR_API void cadr_b200_destroy(cadr_ctx* ctx);
CADR_API int  cadr_b200_device(const cadr_ctx* ctx);            /* CUDA device index, -1 = address-only */
CADR_API int  cadr_b200_sm_count(const cadr_ctx* ctx);
CADR_API cadr_stream cadr_b200_stream(const cadr_ctx* ctx);     /* the context's own stream             */

/* Fence wait — vkWaitForFences in Renderer::executeCopyOperations (Renderer.cpp:982-993).
 * timeout_ns == 0 waits forever; on expiry returns CADR_E_TIMEOUT (CadR::Timeout). */
CADR_API int  cadr_b200_sync(cadr_ctx* ctx, cadr_stream stream, uint64_t timeout_ns);

/* ---- memory ------------------------------------------------------------------------------------- */

/* Device buffer with an address — stands in for vkCreateBuffer + allocatePointerAccessMemory +
 * getBufferDeviceAddress (src/CadR/DataMemory.cpp:36-94, Renderer.cpp:489-591).  256-B aligned. */
CADR_API int  cadr_b200_arena_alloc(cadr_ctx* ctx, size_t bytes, uint64_t* devAddr);
CADR_API int  cadr_b200_arena_free(cadr_ctx* ctx, uint64_t devAddr);

/* Host-visible, host-cached, mapped staging block — StagingMemory::StagingMemory
 * (src/CadR/StagingMemory.cpp:29-76).  Pinned so that uploads are true DMA. */
CADR_API int  cadr_b200_host_alloc(cadr_ctx* ctx, size_t bytes, void** hostPtr);
CADR_API int  cadr_b200_host_free(cadr_ctx* ctx, void* hostPtr);

CADR_API int  cadr_b200_memcpy_h2d(cadr_ctx* ctx, uint64_t dstAddr, const void* src, size_t bytes, cadr_stream stream);
CADR_API int  cadr_b200_memcpy_d2h(cadr_ctx* ctx, void* dst, uint64_t srcAddr, size_t bytes, cadr_stream stream);
CADR_API int  cadr_b200_memset(cadr_ctx* ctx, uint64_t dstAddr, int value, size_t bytes, cadr_stream stream);

/* ---- upload path -------------------------------------------------------------------------------- */

/* DataMemory::recordUploads (src/CadR/DataMemory.cpp:400-446): copy `n` regions from the host staging
 * block `stagingBase` to device addresses.  Regions of 1 MiB or more go out as individual async DMA
 * copies (what vkCmdCopyBuffer does); the others travel with ONE DMA (the span they cover in the staging
 * block when they are dense in it, else a host-packed copy) into a device mirror and are placed by ONE
 * scatter-copy kernel launch.  `stagingBase` may be NULL: srcOffset is then an
 * absolute host address (regions that live in different staging blocks go out in one call). */
CADR_API int  cadr_b200_upload(cadr_ctx* ctx, const cadr_copy_region* regions, uint32_t n,
                               const void* stagingBase, cadr_stream stream);

/* The same upload in two phases, for a renderer that overlaps the PCIe transfer of frame k + 1's uploads with the GPU work
 * of frame k (the reference blocks on a fence per executeCopyOperations, Renderer.cpp:982-993, so its PCIe and GPU phases
 * are serial).  upload_stage moves every region's bytes into a device-side staging slot on `copyStream` and touches no
 * destination - the frame in flight may still read them; upload_commit makes `stream` wait for the staging (an event, not
 * the host) and places the bytes with ONE scatter launch at HBM speed.  The context has two slots: at most two staged
 * uploads may be outstanding.  *ticket == 0 after an empty stage (commit of 0 is a no-op). */
CADR_API int  cadr_b200_upload_stage(cadr_ctx* ctx, const cadr_copy_region* regions, uint32_t n, const void* stagingBase,
                                     cadr_stream copyStream, uint64_t* ticket);
CADR_API int  cadr_b200_upload_commit(cadr_ctx* ctx, uint64_t ticket, cadr_stream stream);

/* Device-side scatter only: staging already resident in HBM at `stagingDevAddr` (srcOffset relative to
 * it).  This is the kernel the HBM roofline is quoted on for the upload path. */
CADR_API int  cadr_b200_scatter_copy(cadr_ctx* ctx, const cadr_copy_region* regions, uint32_t n,
                                     uint64_t stagingDevAddr, cadr_stream stream);

/* HandleTable::set without the 16 KiB whole-leaf re-upload (src/CadR/HandleTable.cpp:58-68,348-378):
 * walks root -> [mid ->] leaf on the device and stores the 8-byte entry in place. */
CADR_API int  cadr_b200_patch_handles(cadr_ctx* ctx, uint64_t handleTableRoot, uint32_t handleLevel,
                                      const cadr_handle_patch* patches, uint32_t n, cadr_stream stream);

/* ---- drawable processing (Tier R: what processDrawables.comp computes) -------------------------- */

/* vkCmdPushConstants + vkCmdDispatch[Base] of processDrawables.comp
 * (src/CadR/Renderer.cpp:669-692; shader main() :92-113).  Same four addresses, same order.
 * numDrawables must be < 2^30 (Renderer.cpp:687).  numDrawables == 0 is a no-op (Renderer.cpp:600-620). */
CADR_API int  cadr_b200_process_drawables(cadr_ctx* ctx, uint64_t handleTableRoot, uint32_t handleLevel,
                                          uint64_t drawableList, uint64_t indirectOut, uint64_t pointersOut,
                                          uint64_t numDrawables, cadr_stream stream);

/* The whole of Renderer::recordDrawableProcessing (Renderer.cpp:598-720): DMA numDrawables*48 bytes
 * from the host staging list to `drawableList` (:635-644), then process them. */
CADR_API int  cadr_b200_record_drawable_processing(cadr_ctx* ctx, const cadr_drawable_gpu_data* hostDrawableList,
                                                   uint64_t handleTableRoot, uint32_t handleLevel,
                                                   uint64_t drawableList, uint64_t indirectOut, uint64_t pointersOut,
                                                   uint64_t numDrawables, cadr_stream stream);

/* ---- Tier X: frustum culling + LOD selection + stream compaction -------------------------------- */

/* No reference counterpart.  Specification: DESIGN.md "Tier X". */
CADR_API int  cadr_b200_cull_compact(cadr_ctx* ctx, const cadr_cull_params* params, cadr_stream stream);

/* Fill bounds[d] for the `count` drawables listed in drawableIndices (device array of uint32_t), or for drawables
 * [0, count) when drawableIndices is 0.  Reads params->indirectData / drawablePointers (Tier R outputs of the current
 * scene state: numMatrices, matrix list address), params->cullData (model-space spheres) and every matrix of the
 * listed drawables once; lists shorter than CADR_CULL_BOUNDS_MIN_LIST get valid = -1 (testing the bound would cost what
 * evaluating them costs).  One warp per drawable. */
CADR_API int  cadr_b200_compute_drawable_bounds(cadr_ctx* ctx, const cadr_cull_params* params, uint64_t boundsOut,
                                                uint64_t drawableIndices, uint32_t count, cadr_stream stream);

/* cadr_b200_process_drawables + cadr_b200_cull_compact in ONE pass over the drawable list: the first kernel also
 * resolves the handles and writes params->indirectData / params->drawablePointers (which are OUTPUTS here, with
 * exactly the contents cadr_b200_process_drawables produces), so the 48-byte records are read once per frame and
 * the 16+32-byte Tier R records are not read back.  Results are identical to the two-call sequence. */
CADR_API int  cadr_b200_process_and_cull(cadr_ctx* ctx, const cadr_cull_params* params, cadr_stream stream);

/* ---- multi-GPU plumbing: one process per GPU, buffers shared through CUDA IPC ------------------------------- */

#define CADR_IPC_HANDLE_BYTES 64
/* Export a buffer returned by cadr_b200_arena_alloc / map a peer's exported buffer into this process. */
CADR_API int  cadr_b200_ipc_export(cadr_ctx* ctx, uint64_t devAddr, unsigned char handle[CADR_IPC_HANDLE_BYTES]);
/* The same for ANY device address inside a cudaMalloc allocation (a slice of a bigger buffer, memory handed out by another
 * allocator of the process): exports the allocation that holds it and reports where devAddr lies inside;
 * the importer adds *offset to the address cadr_b200_ipc_import returns. */
CADR_API int  cadr_b200_ipc_export_range(cadr_ctx* ctx, uint64_t devAddr, unsigned char handle[CADR_IPC_HANDLE_BYTES], uint64_t* offset);
CADR_API int  cadr_b200_ipc_import(cadr_ctx* ctx, const unsigned char handle[CADR_IPC_HANDLE_BYTES], uint64_t* devAddr);
CADR_API int  cadr_b200_ipc_close(cadr_ctx* ctx, uint64_t devAddr);
/* Copy this rank's counters into slot `rank` of every peer's gathered counters, then raise flag[rank] = frameSeq
 * on every peer (release at system scope). */
CADR_API int  cadr_b200_exchange_publish(cadr_ctx* ctx, const cadr_exchange_sync* sync, cadr_stream stream);
/* Both in one launch (one kernel boundary less per frame).  Only where all ranks run concurrently: the kernel spins until
 * every peer has published frameSeq. */
CADR_API int  cadr_b200_exchange_publish_and_wait(cadr_ctx* ctx, const cadr_exchange_sync* sync, cadr_stream stream);
/* Block the stream (not the host) until every peer's flag in the LOCAL flag array has reached frameSeq - or, with
 * sync->timeoutMs > 0, until that budget is spent (both waits; see cadr_exchange_sync). */
CADR_API int  cadr_b200_exchange_wait(cadr_ctx* ctx, const cadr_exchange_sync* sync, cadr_stream stream);

/* Pull every peer's compacted instance-index runs into the local gathered index buffer (stream-ordered; call it after
 * cadr_b200_exchange_wait of the frame).  Inbound NVLink traffic: 4 B x survivors of all other ranks. */
CADR_API int  cadr_b200_exchange_pull_instances(cadr_ctx* ctx, const cadr_exchange_pull* pull, cadr_stream stream);

/* ---- export to a Vulkan consumer or another process (optional in north_star) ------------------- */

/* A device buffer whose memory can be handed out as a POSIX file descriptor (CUDA virtual-memory API: cuMemCreate with
 * an exportable handle type, mapped read/write on the context's device).  Usable wherever an arena_alloc address is
 * (e.g. as cmdOut / instOut / counters of cadr_b200_cull_compact).  allocatedBytes (>= bytes, a multiple of the
 * allocation granularity, 2 MiB on B200) is the size the importer has to state.
 * Counterpart in the reference: the opaque-fd sharing of examples/OpenGLInteroperability/main.cpp:644-651 (export)
 * and :1488-1490 (import), in the other direction: a Vulkan consumer imports the descriptor with
 * VkImportMemoryFdInfoKHR{ handleType = VK_EXTERNAL_MEMORY_HANDLE_TYPE_OPAQUE_FD_BIT, fd } + VkMemoryAllocateInfo{
 * allocationSize = allocatedBytes } and binds a VkBuffer (INDIRECT_BUFFER | STORAGE_BUFFER | SHADER_DEVICE_ADDRESS). */
CADR_API int  cadr_b200_external_alloc(cadr_ctx* ctx, size_t bytes, uint64_t* devAddr, size_t* allocatedBytes);
/* A new descriptor for the buffer; the caller owns it (a successful Vulkan import takes ownership, otherwise close()). */
CADR_API int  cadr_b200_external_export_fd(cadr_ctx* ctx, uint64_t devAddr, int* fd);
/* Map a buffer exported by another context or process (a CUDA consumer; also what the tests use, since no Vulkan ICD
 * exists on the test machines).  The descriptor stays the caller's.  Release with cadr_b200_external_free. */
CADR_API int  cadr_b200_external_import_fd(cadr_ctx* ctx, int fd, size_t allocatedBytes, uint64_t* devAddr);
CADR_API int  cadr_b200_external_free(cadr_ctx* ctx, uint64_t devAddr);

/* Size of the counters buffer for `numStateSets` StateSets. */
CADR_API size_t cadr_b200_cull_counters_bytes(uint32_t numStateSets);

/* ---- consumer-side contract check -------------------------------------------------------------- */

/* Walk drawables [first, first+n) of the Tier R outputs like the reference's vertex shader
 * (examples/RenderingPerformance/shader.vert:99-113): for every instance and vertex of every draw fetch
 * indices[gl_VertexIndex], the 12-byte vertex it selects and matrices[gl_InstanceIndex], and fold them into
 * digestOut[0] (order-independent 64-bit sum of hashes) and digestOut[1] (number of fetches).  digestOut: 16 bytes of
 * device memory.  Geometry must be real (indices < vertex count). */
CADR_API int  cadr_b200_consume_check(cadr_ctx* ctx, uint64_t indirectData, uint64_t drawablePointers, uint64_t firstDrawable,
                                      uint64_t numDrawables, uint64_t digestOut, cadr_stream stream);
/* The same for one draw range of a cadr_b200_cull_compact result, read the way vkCmdDrawIndexedIndirectCount would
 * (count from the counters buffer, at most maxCommands draws; instance k of a command is matrix
 * instOut[firstInstance + k]).  The digest is keyed by drawable index, so it does not depend on emission order or on
 * how a long list was cut into commands.
 * Multi-GPU: to walk rank r's commands on another GPU point cmdOut / ptrOut / tagOut at slot r * exchangeCmdCapacity of
 * the local gathered arrays, counters at rank r's block of the gathered counters, stateSetRegions at a copy of rank r's
 * region table, instOut at rank r's instance-index buffer as mapped here, and set addressDelta (see cadr_cull_params). */
CADR_API int  cadr_b200_consume_check_culled(cadr_ctx* ctx, const cadr_cull_params* params, uint32_t range, uint32_t maxCommands,
                                             uint64_t digestOut, cadr_stream stream);

/* ---- timing (FrameInfo timestamps, src/CadR/Renderer.cpp:436,660,696,778) ----------------------- */

/* When enabled, process_drawables / cull_compact / upload bracket each kernel with CUDA events.
 * cadr_b200_kernel_times returns, after a sync, the milliseconds of the most recent call:
 * [0] process_drawables, [1] cull small-list kernel (thread per drawable), [2] cull list kernel (work items),
 * [3] scatter copy, [4] handle patch.  Unused slots are 0. */
CADR_API int  cadr_b200_set_profiling(cadr_ctx* ctx, int enabled);
CADR_API int  cadr_b200_kernel_times(cadr_ctx* ctx, float* ms, uint32_t n);
/* How many kernels of this library the context has launched since creation. */
CADR_API uint64_t cadr_b200_launch_count(const cadr_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif
