// knn.cu -- mean squared distance to the 3 nearest neighbours (K12-K16) for sm_100a.
//
// Replaces SimpleKNN::knn (src/simple_knn.cu:185-221) behind distCUDA2 (src/spatial.cu:15-27):
// bounding box (min/max seeded with 0 as the reference does, simple_knn.cu:191-199), 30-bit
// Morton codes (:45-70), stable sort of point ids by code, 1024-point boxes (:78-117), and per
// point a 3-NN search over the +-3 Morton neighbours followed by a box-pruned exhaustive scan
// (:147-183).  Differences: no cub/thrust (the library's own onesweep sort), no allocation,
// no host synchronisation -- the bounding box stays on the device.
#include "common.cuh"

#define FLT_MAX_C 3.402823466e+38f

namespace gsb {

constexpr int KNN_BOX = 1024;
constexpr int KNN_THREADS = 256;

struct KnnLayout {
    size_t header, minmax, keys0, keys1, vals0, vals1, hist, lookback, boxes, total;
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
knn_box_kernel(int P, const float* __restrict__ pts, const uint32_t* __restrict__ order, float* __restrict__ boxes)
{
    __shared__ float s_red[6][KNN_BOX / 32];
    const int i = blockIdx.x * KNN_BOX + threadIdx.x;
    float mn[3] = {FLT_MAX_C, FLT_MAX_C, FLT_MAX_C}, mx[3] = {-FLT_MAX_C, -FLT_MAX_C, -FLT_MAX_C};
    if (i < P) {
        const uint32_t id = order[i];
#pragma unroll
        for (int k = 0; k < 3; k++) mn[k] = mx[k] = pts[3 * (size_t)id + k];
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

__device__ __forceinline__ void update_kbest3(float rx, float ry, float rz, const float* __restrict__ q, float* knn)
{
    const float dx = __fsub_rn(q[0], rx), dy = __fsub_rn(q[1], ry), dz = __fsub_rn(q[2], rz);
    float dist = nv3(dx, dx, dy, dy, dz, dz);
#pragma unroll
    for (int j = 0; j < 3; j++)
        if (knn[j] > dist) {
            const float t = knn[j];
            knn[j] = dist;
            dist = t;
        }
}

__global__ void __launch_bounds__(KNN_THREADS)
knn_dist_kernel(int P, const float* __restrict__ pts, const uint32_t* __restrict__ order, const float* __restrict__ boxes,
                int num_boxes, float* __restrict__ out)
{
    const int i = blockIdx.x * KNN_THREADS + threadIdx.x;
    if (i >= P) return;
    const uint32_t id = order[i];
    const float px = pts[3 * (size_t)id], py = pts[3 * (size_t)id + 1], pz = pts[3 * (size_t)id + 2];
    float best[3] = {FLT_MAX_C, FLT_MAX_C, FLT_MAX_C};
    for (int j = max(0, i - 3); j <= min(P - 1, i + 3); j++) {
        if (j == i) continue;
        update_kbest3(px, py, pz, pts + 3 * (size_t)order[j], best);
    }
    const float reject = best[2];
    best[0] = best[1] = best[2] = FLT_MAX_C;
    const float p3[3] = {px, py, pz};
    for (int b = 0; b < num_boxes; b++) {
        float d[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const float lo = boxes[6 * (size_t)b + k], hi = boxes[6 * (size_t)b + 3 + k];
            if (p3[k] < lo || p3[k] > hi) d[k] = fminf(fabsf(__fsub_rn(p3[k], lo)), fabsf(__fsub_rn(p3[k], hi)));
        }
        const float dist = nv3(d[0], d[0], d[1], d[1], d[2], d[2]);
        if (dist > reject || dist > best[2]) continue;
        const int j1 = min(P, (b + 1) * KNN_BOX);
        for (int j = b * KNN_BOX; j < j1; j++) {
            if (j == i) continue;
            update_kbest3(px, py, pz, pts + 3 * (size_t)order[j], best);
        }
    }
    out[id] = __fdiv_rn(__fadd_rn(__fadd_rn(best[0], best[1]), best[2]), 3.0f);
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
    knn_box_kernel<<<L.num_boxes, KNN_BOX, 0, s>>>(P, points, vbuf[0], boxes);
    GSB_LAUNCH_CHECK();
    knn_dist_kernel<<<g, KNN_THREADS, 0, s>>>(P, points, vbuf[0], boxes, L.num_boxes, mean_dist2);
    GSB_LAUNCH_CHECK();
    return GSB_OK;
}

}  // namespace gsb
