// Launcher of the device-side crop extraction (crop_ops.cu; per-pixel arithmetic in crop_math.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "crop_math.h"

namespace ytk {

// pages: [n_pages][H0][W0][3] uint8 BGR on the device; geoms_dev: n_crops CropGeom records on the device;
// scratch / canvases: device buffers the records' roi_off / pix_off point into.  Two launches on `st`.
int launch_extract_crops(const uint8_t* pages, int H0, int W0, const CropGeom* geoms_dev, int n_crops,
                         uint8_t* scratch, uint8_t* canvases, cudaStream_t st);

// src [n][sh][sw][3] -> dst [n][dh][dw][3] = cv2.resize(page, None, fx=0.5, fy=0.5, INTER_AREA); dh / dw = cvRound(sh / 2),
// cvRound(sw / 2) (checked by the caller).
int launch_halve_pages(const uint8_t* src, int n, int sh, int sw, uint8_t* dst, int dh, int dw, cudaStream_t st);

}  // namespace ytk
