// knn.cu -- mean squared distance to the 3 nearest neighbours (K12-K16) for sm_100a.
//
// Replaces SimpleKNN::knn (src/simple_knn.cu:185-221) behind distCUDA2 (src/spatial.cu:15-27):
// bounding box (min/max seeded with 0 as the reference does, simple_knn.cu:191-199), 30-bit
// Morton codes (:45-70), stable sort of point ids by code, 1024-point boxes (:78-117), and per
// point the mean of the three smallest squared distances to other points (:147-183).  That quantity does not depend
// on how the candidates are visited, so only the pruning rules of the reference are kept (3rd-best distance among the
// +-3 Morton neighbours, box distance) and the search itself is organised for the GPU:
//   * the points are gathered ONCE into Morton order as float4 (coalesced 128-bit loads from then on);
//   * a CTA owns 256 consecutive (= spatially coherent) queries, reduces their bounding box and largest rejection
//     radius, and compacts the boxes that can matter to ANY of them into a short shared-memory list (a handful out of
//     ~1000) instead of every thread walking every box;
//   * a candidate box is staged in shared memory with LDG.128 by the whole CTA and scanned by the threads that still
//     need it with broadcast LDS.128 (the classic tiled all-pairs pattern) instead of one scalar gather per pair.
// No cub/thrust (the library's own onesweep sort), no allocation, no host synchronisation -- the bounding box stays on
// the device.
#include "common.cuh"

#define FLT_MAX_C 3.402823466e+38f

namespace gsb {

constexpr int KNN_BOX = 1024;
constexpr int KNN_THREADS = 256;

struct KnnLayout {
    size_t header, minmax, keys0, keys1, vals0, vals1, hist, lookback, boxes, sorted, total;
    int sort_tiles, num_boxes;
    static KnnLayout make(int P)
    {
        KnnLayout L;
        size_t off = 0;
        auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
        const size_t Pz = (size_t)(P > 0 ? P : 1);
        L.sort_tiles = (int)((Pz + SORT_TILE - 1) / SORT_TILE);
        L.num_boxes = (int)((Pz + KNN_BOX - 1) / KNN_BOX);
        L.header = take(sizeof(GeomHeader));
        L.minmax = take(6 * 4);
        L.keys0 = take(Pz * 8); L.keys1 = take(Pz * 8); L.vals0 = take(Pz * 4); L.vals1 = take(Pz * 4);
        L.hist = take((size_t)SORT_MAX_PASSES * SORT_RADIX * 4);
        L.lookback = take((size_t)4 * L.sort_tiles * SORT_RADIX * 4);
        L.boxes = take((size_t)L.num_boxes * 6 * 4);
        L.sorted = take(Pz * 16);   // the points in Morton order, float4
        L.total = off;
        return L;
    }
};

size_t knn_workspace_bytes(int P) { return KnnLayout::make(P).total; }

__device__ __forceinline__ void atomic_max_float(float* addr, float v)
{
    if (v >= 0.f) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
    else atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}
__device__ __forceinline__ void atomic_min_float(float* addr, float v)
{
    if (v >= 0.f) atomicMin(reinterpret_cast<int*>(addr), __float_as_int(v));
    else atomicMax(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

__global__ void __launch_bounds__(KNN_THREADS)
knn_minmax_kernel(int P, const float* __restrict__ pts, float* __restrict__ minmax)
{
    float mn[3] = {0.f, 0.f, 0.f}, mx[3] = {0.f, 0.f, 0.f};  // seeded with 0 (simple_knn.cu:191-199)
    for (int i = blockIdx.x * KNN_THREADS + threadIdx.x; i < P; i += gridDim.x * KNN_THREADS)
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const float v = pts[3 * (size_t)i + k];
            mn[k] = fminf(mn[k], v);
            mx[k] = fmaxf(mx[k], v);
        }
#pragma unroll
    for (int k = 0; k < 3; k++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn[k] = fminf(mn[k], __shfl_xor_sync(0xffffffffu, mn[k], o));
            mx[k] = fmaxf(mx[k], __shfl_xor_sync(0xffffffffu, mx[k], o));
        }
        if (lane_id() == 0) {
            atomic_min_float(minmax + k, mn[k]);
            atomic_max_float(minmax + 3 + k, mx[k]);
        }
    }
}

__device__ __forceinline__ uint32_t prep_morton(uint32_t x)  // simple_knn.cu:45-52
{
    x = (x | (x << 16)) & 0x030000FF;
    x = (x | (x << 8)) & 0x0300F00F;
    x = (x | (x << 4)) & 0x030C30C3;
    x = (x | (x << 2)) & 0x09249249;
    return x;
}

__global__ void __launch_bounds__(KNN_THREADS)
knn_morton_kernel(int P, const float* __restrict__ pts, const float* __restrict__ minmax, uint64_t* __restrict__ keys,
                  uint32_t* __restrict__ vals, GeomHeader* __restrict__ hdr)
{
    const int i = blockIdx.x * KNN_THREADS + threadIdx.x;
    if (i == 0) hdr->num_rendered_clamped = (uint32_t)P;
    if (i >= P) return;
    uint32_t c[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const float f = __fmul_rn(__fdiv_rn(__fsub_rn(pts[3 * (size_t)i + k], minmax[k]), __fsub_rn(minmax[3 + k], minmax[k])),
                                  1023.0f);
        c[k] = prep_morton((uint32_t)f);
    }
    keys[i] = (uint64_t)(c[0] | (c[1] << 1) | (c[2] << 2));
    vals[i] = (uint32_t)i;
}

__global__ void __launch_bounds__(KNN_BOX)
knn_box_kernel(int P, const float* __restrict__ pts, const uint32_t* __restrict__ order, float* __restrict__ boxes,
               float4* __restrict__ sorted)
{
    __shared__ float s_red[6][KNN_BOX / 32];
    const int i = blockIdx.x * KNN_BOX + threadIdx.x;
    float mn[3] = {FLT_MAX_C, FLT_MAX_C, FLT_MAX_C}, mx[3] = {-FLT_MAX_C, -FLT_MAX_C, -FLT_MAX_C};
    if (i < P) {
        const uint32_t id = order[i];
#pragma unroll
        for (int k = 0; k < 3; k++) mn[k] = mx[k] = pts[3 * (size_t)id + k];
        sorted[i] = make_float4(mn[0], mn[1], mn[2], 0.f);
    }
#pragma unroll
    for (int k = 0; k < 3; k++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn[k] = fminf(mn[k], __shfl_xor_sync(0xffffffffu, mn[k], o));
            mx[k] = fmaxf(mx[k], __shfl_xor_sync(0xffffffffu, mx[k], o));
        }
        if (lane_id() == 0) {
            s_red[k][threadIdx.x >> 5] = mn[k];
            s_red[3 + k][threadIdx.x >> 5] = mx[k];
        }
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        float v = s_red[threadIdx.x][0];
        for (int w = 1; w < KNN_BOX / 32; w++) v = threadIdx.x < 3 ? fminf(v, s_red[threadIdx.x][w]) : fmaxf(v, s_red[threadIdx.x][w]);
        boxes[6 * (size_t)blockIdx.x + threadIdx.x] = v;
    }
}

// Insert one squared distance into the ascending triple (a, b, c).
__device__ __forceinline__ void keep_three_smallest(float d, float& a, float& b, float& c)
{
    const float hi0 = fmaxf(a, d);
    a = fminf(a, d);
    const float hi1 = fmaxf(b, hi0);
    b = fminf(b, hi0);
    c = fminf(c, hi1);
}

// squared distance candidate - query, with the contraction nvcc gives the reference's expression (simple_knn.cu:136-139)
__device__ __forceinline__ float dist2(const float4 cand, const float4 q)
{
    const float dx = __fsub_rn(cand.x, q.x), dy = __fsub_rn(cand.y, q.y), dz = __fsub_rn(cand.z, q.z);
    return nv3(dx, dx, dy, dy, dz, dz);
}

// squared distance from a point to an axis-aligned box (0 inside), the reference's pruning measure (simple_knn.cu:119-129)
__device__ __forceinline__ float point_box_dist2(const float4 p, const float* __restrict__ box)
{
    float d[3] = {0.f, 0.f, 0.f};
    const float pk[3] = {p.x, p.y, p.z};
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const float lo = box[k], hi = box[3 + k];
        if (pk[k] < lo || pk[k] > hi) d[k] = fminf(fabsf(__fsub_rn(pk[k], lo)), fabsf(__fsub_rn(pk[k], hi)));
    }
    return nv3(d[0], d[0], d[1], d[1], d[2], d[2]);
}

constexpr int KNN_QUERIES = 256;   // consecutive Morton-sorted queries per CTA

__global__ void __launch_bounds__(KNN_QUERIES)
knn_dist_kernel(int P, const float4* __restrict__ sorted, const uint32_t* __restrict__ order, const float* __restrict__ boxes,
                int num_boxes, float* __restrict__ out)
{
    extern __shared__ __align__(16) unsigned char knn_smem[];
    float4* s_tile = reinterpret_cast<float4*>(knn_smem);                       // one box of points, 16 KB
    int* s_cand = reinterpret_cast<int*>(knn_smem + KNN_BOX * sizeof(float4));  // boxes that can matter to this CTA
    __shared__ float s_red[7][KNN_QUERIES / 32];
    __shared__ float s_q[7];
    __shared__ int s_ncand;
    const int tid = threadIdx.x, i = blockIdx.x * KNN_QUERIES + tid;
    const bool valid = i < P;
    const float4 q = valid ? sorted[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    // rejection radius: third smallest distance among the +-3 neighbours along the Morton curve
    float b0 = FLT_MAX_C, b1 = FLT_MAX_C, b2 = FLT_MAX_C;
    if (valid)
        for (int j = max(0, i - 3); j <= min(P - 1, i + 3); j++)
            if (j != i) keep_three_smallest(dist2(sorted[j], q), b0, b1, b2);
    const float reject = b2;
    // the CTA's query bounding box and its largest rejection radius
    float r[7] = {valid ? q.x : FLT_MAX_C, valid ? q.y : FLT_MAX_C, valid ? q.z : FLT_MAX_C,
                  valid ? q.x : -FLT_MAX_C, valid ? q.y : -FLT_MAX_C, valid ? q.z : -FLT_MAX_C, valid ? reject : 0.f};
#pragma unroll
    for (int k = 0; k < 7; k++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float v = __shfl_xor_sync(0xffffffffu, r[k], o);
            r[k] = k < 3 ? fminf(r[k], v) : fmaxf(r[k], v);
        }
        if (lane_id() == 0) s_red[k][tid >> 5] = r[k];
    }
    if (tid == 0) s_ncand = 0;
    __syncthreads();
    if (tid < 7) {
        float v = s_red[tid][0];
        for (int w = 1; w < KNN_QUERIES / 32; w++) v = tid < 3 ? fminf(v, s_red[tid][w]) : fmaxf(v, s_red[tid][w]);
        s_q[tid] = v;
    }
    __syncthreads();
    // boxes closer to the query box than the largest rejection radius (box-to-box distance never exceeds point-to-box)
    for (int b = tid; b < num_boxes; b += KNN_QUERIES) {
        float g2 = 0.f;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const float gap = fmaxf(0.f, fmaxf(boxes[6 * (size_t)b + k] - s_q[3 + k], s_q[k] - boxes[6 * (size_t)b + 3 + k]));
            g2 = fmaf(gap, gap, g2);
        }
        if (!(g2 > s_q[6] * 1.0001f)) s_cand[atomicAdd(&s_ncand, 1)] = b;
    }
    __syncthreads();
    const int ncand = s_ncand;
    b0 = b1 = b2 = FLT_MAX_C;
    for (int c = 0; c < ncand; c++) {
        const int b = s_cand[c];
        const float bd = valid ? point_box_dist2(q, boxes + 6 * (size_t)b) : FLT_MAX_C;
        const bool need = valid && !(bd > reject || bd > b2);   // the reference's two pruning tests
        if (!__syncthreads_or(need)) continue;
        const int first = b * KNN_BOX, count = min(KNN_BOX, P - first);
        for (int k = tid; k < count; k += KNN_QUERIES) s_tile[k] = sorted[first + k];
        __syncthreads();
        if (need) {
            const int self = i - first;   // position of the query inside this box, if it is one of its points
#pragma unroll 4
            for (int k = 0; k < count; k++)
                if (k != self) keep_three_smallest(dist2(s_tile[k], q), b0, b1, b2);
        }
        __syncthreads();   // everyone is done with the tile before the next box overwrites it
    }
    if (valid) out[order[i]] = __fdiv_rn(__fadd_rn(__fadd_rn(b0, b1), b2), 3.0f);
}

int launch_knn(int P, const float* points, float* mean_dist2, char* ws, cudaStream_t s)
{
    const KnnLayout L = KnnLayout::make(P);
    GeomHeader* hdr = reinterpret_cast<GeomHeader*>(ws + L.header);
    float* minmax = reinterpret_cast<float*>(ws + L.minmax);
    uint64_t* kbuf[2] = {reinterpret_cast<uint64_t*>(ws + L.keys0), reinterpret_cast<uint64_t*>(ws + L.keys1)};
    uint32_t* vbuf[2] = {reinterpret_cast<uint32_t*>(ws + L.vals0), reinterpret_cast<uint32_t*>(ws + L.vals1)};
    uint32_t* hist = reinterpret_cast<uint32_t*>(ws + L.hist);
    uint32_t* lookback = reinterpret_cast<uint32_t*>(ws + L.lookback);
    float* boxes = reinterpret_cast<float*>(ws + L.boxes);
    // header + minmax are adjacent; hist + lookback are adjacent
    GSB_CUDA_CHECK(cudaMemsetAsync(ws + L.header, 0, L.keys0 - L.header, s));
    GSB_CUDA_CHECK(cudaMemsetAsync(hist, 0, L.boxes - L.hist, s));
    const int g = (P + KNN_THREADS - 1) / KNN_THREADS;
    knn_minmax_kernel<<<g < NUM_SMS * 8 ? g : NUM_SMS * 8, KNN_THREADS, 0, s>>>(P, points, minmax);
    GSB_LAUNCH_CHECK();
    knn_morton_kernel<<<g, KNN_THREADS, 0, s>>>(P, points, minmax, kbuf[0], vbuf[0], hdr);
    GSB_LAUNCH_CHECK();
    if (int rc = launch_sort_pairs(hdr, kbuf, vbuf, 0, 4, hist, lookback, L.sort_tiles, s)) return rc;  // 30 bits -> 4 digits
    float4* sorted = reinterpret_cast<float4*>(ws + L.sorted);
    knn_box_kernel<<<L.num_boxes, KNN_BOX, 0, s>>>(P, points, vbuf[0], boxes, sorted);
    GSB_LAUNCH_CHECK();
    const size_t dyn = KNN_BOX * sizeof(float4) + (size_t)L.num_boxes * sizeof(int);   // box tile + candidate list
    if (dyn > 48 * 1024) GSB_SET_ATTR_ONCE(knn_dist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (dyn > 200 * 1024) {
        set_error("knn: %d points need %zu bytes of shared memory for the candidate box list", P, dyn);
        return GSB_ERR_UNSUPPORTED;
    }
    {
        StageTimer _t(ST_OTHER, s);
        knn_dist_kernel<<<(P + KNN_QUERIES - 1) / KNN_QUERIES, KNN_QUERIES, dyn, s>>>(P, sorted, vbuf[0], boxes, L.num_boxes, mean_dist2);
        GSB_LAUNCH_CHECK();
    }
    return GSB_OK;
}

}  // namespace gsb
