// jpeg.h -- device-side baseline JPEG writer (encode half): launchers (jpeg.cu) and host helpers
// (jpeg_host.cpp).  See jpeg_core.h for what it reproduces.
#pragma once
#include <stddef.h>
#include "jpeg_core.h"

#if defined(__CUDACC__) || defined(IPG_WITH_CUDA_RUNTIME)
#include <cuda_runtime.h>
#endif

namespace ipg {

enum {
    JPEG_HDR_MAX = 640,      // SOI + DQT + SOF0 + DHT + SOS header = 589 bytes
    JPEG_CHUNK = 512,        // unstuffed scan bytes per stuffing chunk (one warp, 16 bytes per lane)
    JPEG_DCT_MCUS = 32,      // MCUs per k_jpeg_dct / k_jpeg_emit CTA (6 warps: one per block of the MCU); = JPEG_SLOT_STRIDE
    JPEG_SCAN_THREADS = 1024,
    JPEG_STUFF_THREADS = 256,
    JPEG_STUFF_PARTS = 512,  // at most this many k_jpeg_zero / k_jpeg_ffcount / k_jpeg_write CTAs per job (each strides over the job's chunks)
};

struct JpegDctItem { int32_t job; int32_t mcu0; };   // JPEG_DCT_MCUS MCUs starting at mcu0 (k_jpeg_dct and k_jpeg_emit)
struct JpegStuffItem { int32_t job; int16_t part, nparts; }; // CTA `part` of the job's `nparts`

// host helpers (jpeg_host.cpp)
void jpeg_build_tables(int quality, JpegTables *t);
size_t jpeg_build_header(int quality, int w, int h, uint8_t *out); // <= JPEG_HDR_MAX bytes

#if defined(__CUDACC__) || defined(IPG_WITH_CUDA_RUNTIME)
// The whole writer for every job of a batch, in stream order after the kernels that produce the RGBA results:
// k_jpeg_dct -> k_jpeg_offsets -> k_jpeg_zero -> k_jpeg_emit -> k_jpeg_ffcount -> k_jpeg_chunks -> k_jpeg_write.
// JpegJob::result[0..3] must be zero when it starts.
enum { JPEG_LAUNCHES = 7 };
cudaError_t launch_jpeg(const JpegJob *jobs, int n_jobs, const JpegDctItem *dct_items, int n_dct, const JpegStuffItem *stuff_items,
                        int n_stuff, cudaStream_t st);
#endif

} // namespace ipg
