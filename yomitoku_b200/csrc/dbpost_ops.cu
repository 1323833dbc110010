// Device-side front half of the DBNet post-processing (reference postprocessor/dbnet_postporcessor.py:39-82): from the
// probability map to what the per-box geometry on the host needs - a few KB per page instead of the 7.6 MB map.
//
//   bitmap   = prob > thresh                                              (:29-30, :39-43)
//   contours = cv2.findContours(bitmap, RETR_LIST, CHAIN_APPROX_SIMPLE)   (:45-47) -> 8-connected components
//   per contour: minAreaRect(contour) (:100-124) and box_score_fast = mean of prob over fillPoly(contour) (:126-138)
//
// What the host needs from a component is (a) its convex hull - minAreaRect of the contour and of the end points of the
// component's row runs are the same rectangle, bit for bit (same hull vertex set, OpenCV sorts the points before building
// the hull), (b) sum and count of prob over the filled contour, which for a component without holes is the component
// itself, and (c) OpenCV's contour order, which for outer contours is descending raster index of the component's first
// pixel.  All three were checked against OpenCV on random maps (tests/test_host_logic.py, tests/test_gpu_dbpost.py).
// Components with holes produce extra (hole) contours in OpenCV; the kernels count holes with the Euler number
// (#holes = #components - E8) and the caller falls back to the host path for such a page, so the result is exact always.
//
// Kernels: label_seed / label_merge (union-find CCL over 32-pixel chunk runs, root = smallest raster index = first
// pixel), page_stats (components, 2x2 bit-quad counts for the Euler number), emit_runs (one record per row run: root,
// row, first / last column, sum of prob in fp64).  HBM-bound byte work: map + labels are read ~4 times (30 MB per page).
#include "dbpost_ops.h"

#include "gemm_tc.h"

namespace ytk {

namespace {

__device__ __forceinline__ int uf_find(const volatile int* lab, int x) {   // volatile: other threads hook roots concurrently
    int p = lab[x];
    while (p != x) {
        x = p;
        p = lab[x];
    }
    return x;
}

__device__ __forceinline__ void uf_union(int* lab, int a, int b) {
    // roots are the smallest raster index of their set: hang the larger root under the smaller one
    for (;;) {
        a = uf_find(lab, a);
        b = uf_find(lab, b);
        if (a == b) return;
        if (a > b) {
            const int t = a;
            a = b;
            b = t;
        }
        const int old = atomicMin(&lab[b], a);
        if (old == b) return;
        b = old;
    }
}

// labels are PER PAGE raster indices (page offset subtracted) so that a root is the component's first pixel.
// Seed: a warp looks at 32 consecutive pixels; every foreground pixel starts as a child of the first pixel of its row run
// INSIDE that 32-pixel chunk (one ballot, no memory traffic), so the merge pass only has to join chunk runs with each
// other - a text line blob of 400 x 25 pixels costs ~400 unions instead of 10 000 long pointer chases.
__global__ void label_seed_kernel(const float* __restrict__ prob, float thresh, int* __restrict__ lab, int HW, int W,
                                  long long total) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool in = i < total;
    const bool fg = in && prob[i] > thresh;
    const unsigned mask = __ballot_sync(0xffffffffu, fg);
    if (!in) return;
    const int lane = threadIdx.x & 31;
    const int p = (int)(i % HW);
    const int x = p % W;
    const unsigned below = ~mask & ((1u << lane) - 1u);            // background lanes in front of this one
    int start = below ? 32 - __clz(below) : 0;                      // first lane of this lane's run in the chunk
    const int row_first = lane - (lane < x ? lane : x);             // lanes before it belong to the previous row (or page)
    if (start < row_first) start = row_first;
    lab[i] = fg ? p - (lane - start) : -1;
}

__global__ void label_merge_kernel(int* __restrict__ lab, int H, int W, long long total) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int HW = H * W;
    int* page = lab + (i / HW) * (long long)HW;
    const int p = (int)(i % HW);
    if (page[p] < 0) return;
    const int y = p / W, x = p - y * W;
    const bool left = x > 0 && page[p - 1] >= 0;
    // a run that continues across a chunk border: join the two chunk runs
    if (left && (threadIdx.x & 31) == 0) uf_union(page, p, p - 1);
    if (y == 0) return;
    // 8-connectivity with the row above; unions that follow from row adjacency of already joined pixels are skipped
    const bool up = page[p - W] >= 0;
    const bool ul = x > 0 && page[p - W - 1] >= 0;
    const bool ur = x + 1 < W && page[p - W + 1] >= 0;
    if (up) {
        if (!(left && ul)) uf_union(page, p, p - W);               // else: p - left - ul - up are joined already
    } else {
        if (ul && !left) uf_union(page, p, p - W - 1);             // else: left joins ul (it is left's `up`)
        if (ur) uf_union(page, p, p - W + 1);
    }
}

// meta[page] = {runs, components, 4 * E8 accumulator (Q1 - Q3 - 2 QD), overflow}
__global__ void page_stats_kernel(const int* __restrict__ lab, int H, int W, int* __restrict__ meta, long long total_q) {
    // one thread per 2x2 bit quad; quads cover the map extended by one background pixel on every side
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total_q) return;
    const int QW = W + 1, QH = H + 1;
    const int page = (int)(i / ((long long)QW * QH));
    const int q = (int)(i % ((long long)QW * QH));
    const int qy = q / QW, qx = q - qy * QW;           // quad with bottom-right pixel (qy, qx), top-left (qy-1, qx-1)
    const int* pg = lab + (long long)page * H * W;
    auto fg = [&](int y, int x) { return y >= 0 && y < H && x >= 0 && x < W && pg[y * W + x] >= 0; };
    const int a = fg(qy - 1, qx - 1), b = fg(qy - 1, qx), c = fg(qy, qx - 1), d = fg(qy, qx);
    const int n = a + b + c + d;
    int e = 0;
    if (n == 1) e = 1;
    else if (n == 3) e = -1;
    else if (n == 2 && a == d) e = -2;                  // the two diagonal patterns
    if (e != 0) atomicAdd(&meta[page * 4 + 2], e);
    if (d && pg[qy * W + qx] == qy * W + qx) atomicAdd(&meta[page * 4 + 1], 1);   // a root = one component
}

__global__ void emit_runs_kernel(const float* __restrict__ prob, const int* __restrict__ lab, int H, int W,
                                 DbRun* __restrict__ runs, int max_runs, int* __restrict__ meta, long long total) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int HW = H * W;
    const int page = (int)(i / HW);
    const int p = (int)(i % HW);
    const int* pg = lab + (long long)page * HW;
    if (pg[p] < 0) return;
    const int y = p / W, x = p - y * W;
    if (x > 0 && pg[p - 1] >= 0) return;                // not the first pixel of its run
    const float* pr = prob + (long long)page * HW + p;
    double s = 0.0;
    int x1 = x;
    while (x1 < W && pg[p + (x1 - x)] >= 0) {
        s += (double)pr[x1 - x];
        ++x1;
    }
    const int k = atomicAdd(&meta[page * 4 + 0], 1);
    if (k >= max_runs) {
        meta[page * 4 + 3] = 1;
        return;
    }
    DbRun r;
    r.root = uf_find(pg, p);                              // labels are never flattened: only run starts need their root
    r.y = y;
    r.x0 = x;
    r.x1 = x1 - 1;
    r.sum = s;
    runs[(long long)page * max_runs + k] = r;
}

}  // namespace

long long dbpost_scratch_bytes(int n_pages, int H, int W) { return (long long)n_pages * H * W * (long long)sizeof(int); }

int launch_dbpost_front(const float* prob, int n_pages, int H, int W, float thresh, int* labels, DbRun* runs,
                        int max_runs, int* meta, cudaStream_t st) {
    const long long total = (long long)n_pages * H * W;
    if (total <= 0) return 0;
    const int threads = 256;
    const unsigned grid = (unsigned)((total + threads - 1) / threads);
    if (cudaMemsetAsync(meta, 0, sizeof(int) * 4 * n_pages, st) != cudaSuccess) return 1;
    label_seed_kernel<<<grid, threads, 0, st>>>(prob, thresh, labels, H * W, W, total);
    label_merge_kernel<<<grid, threads, 0, st>>>(labels, H, W, total);
    const long long total_q = (long long)n_pages * (H + 1) * (W + 1);
    page_stats_kernel<<<(unsigned)((total_q + threads - 1) / threads), threads, 0, st>>>(labels, H, W, meta, total_q);
    emit_runs_kernel<<<grid, threads, 0, st>>>(prob, labels, H, W, runs, max_runs, meta, total);
    count_launch(4);
    return cudaGetLastError() != cudaSuccess;
}

}  // namespace ytk
