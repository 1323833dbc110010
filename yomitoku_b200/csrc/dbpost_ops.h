// Device-side front half of the DBNet post-processing (dbpost_ops.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ytk {

// One horizontal run of an 8-connected component of (prob > thresh): same layout as ytk_db_run (yomitoku_b200.h).
struct DbRun {
    int32_t root;   // raster index (y * W + x) of the component's first pixel = its id, and its rank in OpenCV's order
    int32_t y;
    int32_t x0;     // first column
    int32_t x1;     // last column (inclusive)
    double sum;     // sum of prob over the run (fp64)
};

long long dbpost_scratch_bytes(int n_pages, int H, int W);
// prob: [n_pages, H, W] fp32 on the device; labels: scratch of dbpost_scratch_bytes(); runs: [n_pages, max_runs];
// meta: [n_pages, 4] int32 = {runs found (may exceed max_runs), components, 4 * Euler number (8-connectivity), overflow}.
int launch_dbpost_front(const float* prob, int n_pages, int H, int W, float thresh, int* labels, DbRun* runs,
                        int max_runs, int* meta, cudaStream_t st);

}  // namespace ytk
