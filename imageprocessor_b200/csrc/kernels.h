// kernels.h -- host-callable launchers of the sm_100a kernels (kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include "ipg_device.h"

namespace ipg {

// Streaming fp32 resample (+ fused watermark copy/blend): one CTA per StreamItem.
// Source rows must be 16-byte aligned (base and stride): they are moved by TMA bulk copies.
cudaError_t launch_stream(const StreamJob *jobs, const StreamItem *items, int n_items,
                          int max_targets, bool any_wm, FixList fix, cudaStream_t st);

// A lean single-target instantiation (k_stream<1, WM, kind>) over jobs with fast_path == kind.
cudaError_t launch_stream_fast(const StreamJob *jobs, const StreamItem *items, int n_items, int kind, bool any_wm,
                               FixList fix, cudaStream_t st);

// The lean streaming resample for planar YCbCr sources (one target per job, cached pass forms only).
cudaError_t launch_stream_planar(const StreamJob *jobs, const StreamItem *items, int n_items, bool nrgba, FixList fix, cudaStream_t st);

// Small-support fp32 resample (vertical upscales, mild downscales), one CTA per 32x8 output tile; flags into `fix`.
cudaError_t launch_direct(const DirectJob *jobs, const DirectItem *items, int n_items, FixList fix, cudaStream_t st);

// fp64 reference-order resample of whole outputs: one CTA per 32x8 output tile.
cudaError_t launch_exact_tiles(const ExactJob *jobs, const ExactItem *items, int n_items,
                               cudaStream_t st);

// fp64 re-evaluation of the pixels a stream kernel flagged (count read on device).
cudaError_t launch_exact_fix(const ExactJob *jobs, int n_jobs, FixList fix, cudaStream_t st);

// Ordered glyph blend in place over each watermark's glyph box (after the copy/convert).
cudaError_t launch_blend(const WatermarkD *wms, const BlendItem *items, int n_items, cudaStream_t st);

// RGBA8 results -> planar YCbCr 4:2:0 as Go's image/jpeg writer derives it (one CTA per 256 x 16 pixels).
cudaError_t launch_rgba_to_ycbcr420(const YccJob *jobs, const YccItem *items, int n_items, cudaStream_t st);

// Patch-only watermark of RGBA8 sources: source box -> blend -> patch buffer / caller frame.
cudaError_t launch_blend_patch(const PatchJob *jobs, const BlendItem *items, int n_items, cudaStream_t st);

// draw.Draw(Src) convert/copy for any layout (glyphs follow in launch_blend).
cudaError_t launch_watermark(const WmJob *jobs, const WmItem *items, int n_items, cudaStream_t st);

int stream_smem_bytes();

} // namespace ipg
