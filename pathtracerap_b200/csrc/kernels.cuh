// kernels.cuh - launch wrappers of the wavefront kernels (one translation unit per kernel family).
#pragma once
#include "device_types.h"
#include "exact_math.cuh"
#include "../../include/ptap.h"

namespace ptap {

constexpr int kTraceBlock = 128;
constexpr int kShadeBlock = 128;
constexpr int kShadeTile = 32;       // one compaction tile = one warp = 32 consecutive slots
constexpr int kScanBlock = 256, kScanSlots = 2048, kScanTiles = kScanSlots / kShadeTile;   // k_scan: slots per CTA, tiles per CTA
constexpr int kGenBlock = 256;
constexpr int kTraceBatch = 32;      // rays a warp takes from the work-stealing cursor per atomic
constexpr int kVoteTri = 8, kVoteInst = 4, kVoteRefill = 4;   // state-machine thresholds of k_trace_bvh (trace_bvh.cu)

// closest hit, grid-compat (trace_grid.cu) and BVH (trace_bvh.cu).  n_fixed < 0: read the count from st->n_active[round].
void launchTraceGrid(const SceneDev& sc, const float4* O, const float4* D, float4* hit, float2* uv, int4* counts, bool count_totals,
                     FrameState* st, int round, int n_fixed, int grid, cudaStream_t stream);
int traceGridOccupancy();
void launchTraceBvh(const SceneDev& sc, const float4* O, const float4* D, float4* hit, float2* uv, int4* counts, bool count_totals,
                    FrameState* st, int round, int n_fixed, int grid, cudaStream_t stream);
int traceBvhOccupancy();

// wavefront.cu
void launchGenerate(const WaveDev& wv, int grid, cudaStream_t stream);
void launchScan(const SceneDev& sc, const WaveDev& wv, int round, const float4* hit, int remaining, int n_fixed, cudaStream_t stream);
void launchShade(const SceneDev& sc, const WaveDev& wv, int round, int in_buf, const float4* hit, int remaining, int n_fixed,
                 int iter_fixed, int* slot_pos, int grid, cudaStream_t stream);
int shadeOccupancy();
void launchResolveHits(const SceneDev& sc, const float4* hit, const float2* uv, int n, PtapHit* out, cudaStream_t stream);
void launchSetIter(FrameState* st, int iter, cudaStream_t stream);
void launchFilmAdd(float* film, const float* add, size_t n, cudaStream_t stream);
void launchExtractNormals(const TriRec* tris, int n, float4* normals, cudaStream_t stream);
void launchGatherTris(const TriRec* tris, const int* tri_id, int n, TriRec* out, cudaStream_t stream);

}  // namespace ptap
