// jpeg.cu -- device-side baseline JPEG writer for the RGBA results of the raster kernels (sm_100a).
//
// SURVEY 8f-3, encode half: the reference ends every operation with jpeg.Encode(buf, img, &jpeg.Options{Quality: 85})
// on the host (operations/resize.go:78-91, thumbnail.go, watermark.go:66-79).  With ipg_op.dst_layout = IPG_LAYOUT_JPEG
// the file is produced here instead, byte for byte what Go's writer emits (jpeg_core.h says how), and only the file
// crosses PCIe: ~0.2-1 instead of 4 bytes per result pixel, and no host encode.
//
// Huffman coding is sequential in the bit position, so the writer is seven data-parallel passes:
//   k_jpeg_dct      one thread per 8 x 8 block: colour transform, FDCT, quantisation; the block's AC coefficients are
//                   entropy-coded on the spot into a private slot (only the DC code depends on another block)
//   k_jpeg_offsets  one CTA per job: bits per MCU (DC deltas need the neighbour's dc) -> exclusive scan = bit offsets
//   k_jpeg_zero     clears exactly the words the scan will occupy (its size is known on the device only)
//   k_jpeg_emit     one thread per block: its DC code, then its slot shift-copied into the MSB-first scan at its bit offset
//   k_jpeg_ffcount  one warp per 512 scan bytes: how many 0xff bytes (each gets a 0x00 stuffed after it)
//   k_jpeg_chunks   one CTA per job: exclusive scan of those counts, file length
//   k_jpeg_write    one warp per 512 scan bytes: header | stuffed scan | EOI
// Integer work, bound by issue slots and L2 latency, far below HBM; it runs in the shadow of the next batch's uploads.
#include "jpeg.h"

namespace ipg {

namespace {

__device__ __forceinline__ uint32_t warp_sum(uint32_t v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ uint32_t warp_inclusive(uint32_t v, int lane)
{
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t n = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += n;
    }
    return v;
}
// Exclusive prefix of one value per thread over a 1024-thread CTA; returns the CTA total in `total`.
__device__ __forceinline__ uint32_t block_exclusive(uint32_t v, uint32_t *smem32, uint32_t &total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t inc = warp_inclusive(v, lane);
    if (lane == 31) smem32[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        const uint32_t w = smem32[lane];
        const uint32_t winc = warp_inclusive(w, lane);
        smem32[lane] = winc - w;
        if (lane == 31) smem32[32] = winc;
    }
    __syncthreads();
    const uint32_t r = smem32[warp] + inc - v;
    total = smem32[32];
    __syncthreads(); // smem32 is reused by the next round
    return r;
}

// ---- k_jpeg_dct ---------------------------------------------------------------------------------------------------
// A CTA takes JPEG_DCT_MCUS = 32 consecutive MCUs (scan order) of one job; they may wrap to the next MCU row, every MCU
// is staged on its own.  The 16 x 16 pixels of each MCU are staged in shared memory with the writer's edge clamp
// applied; MCU m's rows sit 17 words apart from MCU m+1's so that the 32 lanes of a warp (one MCU each) hit 32
// different banks.  Warps 0-3 convert the luma quadrants of the 32 MCUs (and leave the 2 x 2 chroma means of their pixels
// in shared memory, which is how writer.go's rgbaToYCbCr + scale split the work too); then warp b runs FDCT,
// quantisation and AC coding of block b (Y0..Y3, Cb, Cr).
enum { DCT_ROW_WORDS = JPEG_DCT_MCUS * 17 };
struct DctSmem {
    union {
        uint32_t px[16 * DCT_ROW_WORDS];   // the staged pixels: dead once every luma warp has converted its quadrants
        int16_t coef[64][192];             // then: non-zero quantised AC coefficients, [zig][thread] (lanes side by side; the coding threads are 0..191)
    };
    uint32_t cmean[2][JPEG_DCT_MCUS * 17]; // Cb / Cr: the 16 words (8 x 8 means) of each MCU, MCUs 17 words apart
    int32_t half[2][64];
    uint32_t recip[2][64];
    uint32_t lut_ac[2][256];
};
struct CoefSmem {
    int16_t *p; // this thread's column
    __device__ __forceinline__ void set(int zig, int v) { p[zig * 192] = (int16_t)v; }
    __device__ __forceinline__ int get(int zig) const { return p[zig * 192]; }
};
// 4 CTAs per SM (80 registers, a 24-byte spill) beat 3 (96 registers, none) by 6 %: the kernel is latency-bound at 18 warps.
// Measured and dropped: 8 warps with the chroma means on warps 4-7 beside the luma warps (95-111 us against 95).
#ifndef IPG_JPEG_DCT_CTAS
#define IPG_JPEG_DCT_CTAS 4
#endif
#define IPG_JPEG_DCT_WARPS 6
enum { DCT_THREADS = IPG_JPEG_DCT_WARPS * 32 };
__global__ void __launch_bounds__(DCT_THREADS, IPG_JPEG_DCT_CTAS) k_jpeg_dct(const JpegJob *__restrict__ jobs, const JpegDctItem *__restrict__ items)
{
    __shared__ DctSmem sm;
    const JpegDctItem it = items[blockIdx.x];
    const JpegJob &J = jobs[it.job];
    const JpegTables &T = *J.tab;
    const int xmax = J.w - 1, ymax = J.h - 1;
    const int n_here = min((int)JPEG_DCT_MCUS, J.n_mcu - it.mcu0);
    for (int k = threadIdx.x; k < 128; k += blockDim.x) {
        sm.half[k >> 6][k & 63] = T.half[k >> 6][k & 63];
        sm.recip[k >> 6][k & 63] = T.recip[k >> 6][k & 63];
    }
    for (int k = threadIdx.x; k < 512; k += blockDim.x) sm.lut_ac[k >> 8][k & 255] = T.lut[(k >> 8) * 2 + 1][k & 255];
    // stage: 4 pixels per piece (one 16-byte load when they are all inside the image); consecutive threads take
    // consecutive 16-byte pieces of one image row, MCU after MCU
    const int mx0 = it.mcu0 % J.mcu_w, my0 = it.mcu0 / J.mcu_w;
    constexpr int kPieces = 16 * JPEG_DCT_MCUS * 4, kRounds = (kPieces + DCT_THREADS - 1) / DCT_THREADS; // 2048 pieces, 11 per thread
    uint4 v[kRounds];
#pragma unroll
    for (int r = 0; r < kRounds; r++) { // all the loads first: 11 in flight per thread
        const int q = threadIdx.x + r * DCT_THREADS;
        const int row = q >> 7, m = (q >> 2) & 31, quad = q & 3;
        v[r] = make_uint4(0u, 0u, 0u, 0u);
        if (q >= kPieces || m >= n_here) continue;
        int mx = mx0 + m, my = my0;
        while (mx >= J.mcu_w) { mx -= J.mcu_w; my++; }
        const int x0 = mx * 16 + quad * 4, y = min(my * 16 + row, ymax);
        const uint8_t *rowp = J.rgba + (size_t)y * (size_t)J.rgba_pitch;
        if (x0 + 3 <= xmax) {
            v[r] = __ldg(reinterpret_cast<const uint4 *>(rowp + (size_t)x0 * 4));
        } else {
            v[r].x = __ldg(reinterpret_cast<const uint32_t *>(rowp + (size_t)min(x0, xmax) * 4));
            v[r].y = __ldg(reinterpret_cast<const uint32_t *>(rowp + (size_t)min(x0 + 1, xmax) * 4));
            v[r].z = __ldg(reinterpret_cast<const uint32_t *>(rowp + (size_t)min(x0 + 2, xmax) * 4));
            v[r].w = __ldg(reinterpret_cast<const uint32_t *>(rowp + (size_t)min(x0 + 3, xmax) * 4));
        }
    }
#pragma unroll
    for (int r = 0; r < kRounds; r++) {
        const int q = threadIdx.x + r * DCT_THREADS;
        if (q >= kPieces) continue;
        const int row = q >> 7, m = (q >> 2) & 31, quad = q & 3;
        uint32_t *d = sm.px + row * DCT_ROW_WORDS + m * 17 + quad * 4;
        d[0] = v[r].x; d[1] = v[r].y; d[2] = v[r].z; d[3] = v[r].w;
    }
    __syncthreads();
    const int wrp = threadIdx.x >> 5, m = threadIdx.x & 31;
    const bool active = m < n_here;
    int32_t b[64];
    {
        const uint32_t *base = sm.px + m * 17;
        auto px = [base](int lx, int ly) { return base[ly * DCT_ROW_WORDS + lx]; };
        if (wrp < 4 && active) { // the four luma warps convert their quadrants and leave the chroma means for warps 4 and 5
            uint32_t cbw[4], crw[4];
            jpeg_quadrant(wrp, b, cbw, crw, px);
#pragma unroll
            for (int r = 0; r < 4; r++) {
                sm.cmean[0][m * 17 + jpeg_chroma_word(wrp, r)] = cbw[r];
                sm.cmean[1][m * 17 + jpeg_chroma_word(wrp, r)] = crw[r];
            }
        }
    }
    __syncthreads();
    const int blk = wrp; // warp b codes block b
    if (!active) return;
    const int mcu = it.mcu0 + m, q = blk < 4 ? 0 : 1;
    if (blk >= 4) {
        const uint32_t *cw = sm.cmean[blk - 4] + m * 17;
        jpeg_chroma_block(b, [cw](int k) { return cw[k]; });
    }
    jpeg_block_code(b, sm.half[q], sm.recip[q], sm.lut_ac[q], J.acs + jpeg_slot_index(mcu, blk), JPEG_SLOT_STRIDE, J.side + (size_t)mcu * 6 + blk,
                    CoefSmem{sm.coef[0] + threadIdx.x});
}

// ---- k_jpeg_offsets -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(JPEG_SCAN_THREADS) k_jpeg_offsets(const JpegJob *__restrict__ jobs)
{
    __shared__ uint32_t sm[33];
    const JpegJob &J = jobs[blockIdx.x];
    const JpegTables &T = *J.tab;
    uint32_t carry = 0;
    for (int base = 0; base < J.n_mcu; base += JPEG_SCAN_THREADS * 4) {
        const int m0 = base + threadIdx.x * 4;
        uint32_t b[4], s = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            b[k] = m0 + k < J.n_mcu ? jpeg_mcu_bits(T, J.side, m0 + k) : 0u;
            s += b[k];
        }
        uint32_t total;
        uint32_t off = carry + block_exclusive(s, sm, total);
#pragma unroll
        for (int k = 0; k < 4; k++) {
            if (m0 + k < J.n_mcu) J.mcu_off[m0 + k] = off;
            off += b[k];
        }
        if (total > 0xffffffffu - carry) carry = 0xffffffffu; else carry += total; // saturate: reported as "does not fit"
    }
    if (threadIdx.x == 0) {
        const uint64_t bytes = ((uint64_t)carry + 7) >> 3;
        J.result[3] = carry;
        J.result[2] = (uint32_t)bytes;
        if (carry == 0xffffffffu || bytes > J.cap_bytes) J.result[1] = 1;
    }
}

// ---- k_jpeg_zero ----------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(JPEG_STUFF_THREADS) k_jpeg_zero(const JpegJob *__restrict__ jobs, const JpegStuffItem *__restrict__ items)
{
    const JpegStuffItem it = items[blockIdx.x];
    const JpegJob &J = jobs[it.job];
    if (J.result[1]) return;
    const uint32_t n16 = (J.result[2] + JPEG_CHUNK - 1) / JPEG_CHUNK * (JPEG_CHUNK / 16); // whole chunks: the stuffing passes read them
    uint4 *w = reinterpret_cast<uint4 *>(J.words);
    for (uint32_t k = it.part * JPEG_STUFF_THREADS + threadIdx.x; k < n16; k += it.nparts * JPEG_STUFF_THREADS)
        w[k] = make_uint4(0u, 0u, 0u, 0u);
}

// ---- k_jpeg_emit ----------------------------------------------------------------------------------------------------
// Same CTA shape as k_jpeg_dct: 32 MCUs, warp b = block b.  A block's place in the scan is its MCU's offset plus the
// bits of the blocks before it in the MCU (their DC codes depend on their predictors: the side words of the 32 MCUs
// and of the one before them are staged in shared memory).
__global__ void __launch_bounds__(192) k_jpeg_emit(const JpegJob *__restrict__ jobs, const JpegDctItem *__restrict__ items)
{
    __shared__ uint32_t side_s[(JPEG_DCT_MCUS + 1) * 6];
    __shared__ uint32_t lut_dc[2][16];
    const JpegDctItem it = items[blockIdx.x];
    const JpegJob &J = jobs[it.job];
    if (J.result[1]) return;
    const int n_here = min((int)JPEG_DCT_MCUS, J.n_mcu - it.mcu0);
    for (int k = threadIdx.x; k < (n_here + 1) * 6; k += blockDim.x) {
        const int g = (it.mcu0 - 1) * 6 + k;
        side_s[k] = g >= 0 ? J.side[g] : 0u;
    }
    if (threadIdx.x < 32) lut_dc[threadIdx.x >> 4][threadIdx.x & 15] = J.tab->lut[(threadIdx.x >> 4) * 2][threadIdx.x & 15];
    __syncthreads();
    const int blk = threadIdx.x >> 5, m = threadIdx.x & 31;
    if (m >= n_here) return;
    const int mcu = it.mcu0 + m, mi = m + 1;
    const bool first = mcu == 0;
    uint32_t off = J.mcu_off[mcu];
    for (int b = 0; b < blk; b++) off += jpeg_block_bits(lut_dc[0], lut_dc[1], side_s, mi, b, first);
    jpeg_block_emit(J.words, lut_dc[blk < 4 ? 0 : 1], side_s[mi * 6 + blk], jpeg_prev_dc(side_s, mi, blk, first), off,
                    J.acs + jpeg_slot_index(mcu, blk), JPEG_SLOT_STRIDE, mcu == J.n_mcu - 1 && blk == 5);
}

// ---- byte stuffing --------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t ff_bytes(uint32_t w) // bytes of w equal to 0xff
{
    uint32_t t = w & (w >> 4); // bit 0 of each byte ends up as the AND of its eight bits (no term crosses a byte)
    t &= t >> 2;
    t &= t >> 1;
    return __popc(t & 0x01010101u);
}

__global__ void __launch_bounds__(JPEG_STUFF_THREADS) k_jpeg_ffcount(const JpegJob *__restrict__ jobs, const JpegStuffItem *__restrict__ items)
{
    const JpegStuffItem it = items[blockIdx.x];
    const JpegJob &J = jobs[it.job];
    if (J.result[1]) return;
    const uint32_t U = J.result[2], n_chunks = (U + JPEG_CHUNK - 1) / JPEG_CHUNK;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint32_t c = it.part * (JPEG_STUFF_THREADS / 32) + warp; c < n_chunks; c += it.nparts * (JPEG_STUFF_THREADS / 32)) {
        const uint4 v = *reinterpret_cast<const uint4 *>(J.words + (size_t)c * (JPEG_CHUNK / 4) + lane * 4); // zero past the scan
        const uint32_t n = warp_sum(ff_bytes(v.x) + ff_bytes(v.y) + ff_bytes(v.z) + ff_bytes(v.w));
        if (lane == 0) J.chunk_off[c] = n;
    }
}

__global__ void __launch_bounds__(JPEG_SCAN_THREADS) k_jpeg_chunks(const JpegJob *__restrict__ jobs)
{
    __shared__ uint32_t sm[33];
    const JpegJob &J = jobs[blockIdx.x];
    if (J.result[1]) return;
    const uint32_t U = J.result[2], n_chunks = (U + JPEG_CHUNK - 1) / JPEG_CHUNK;
    uint32_t carry = 0;
    for (uint32_t base = 0; base < n_chunks; base += JPEG_SCAN_THREADS) {
        const uint32_t c = base + threadIdx.x;
        const uint32_t v = c < n_chunks ? J.chunk_off[c] : 0u;
        uint32_t total;
        const uint32_t off = carry + block_exclusive(v, sm, total);
        if (c < n_chunks) J.chunk_off[c] = off;
        carry += total;
    }
    if (threadIdx.x == 0) {
        const uint64_t len = (uint64_t)J.hdr_len + U + carry + 2;
        if (len > J.out_cap) J.result[1] = 1;
        else J.result[0] = (uint32_t)len;
    }
}

__global__ void __launch_bounds__(JPEG_STUFF_THREADS) k_jpeg_write(const JpegJob *__restrict__ jobs, const JpegStuffItem *__restrict__ items)
{
    const JpegStuffItem it = items[blockIdx.x];
    const JpegJob &J = jobs[it.job];
    if (J.result[1]) return;
    const uint32_t U = J.result[2], n_chunks = (U + JPEG_CHUNK - 1) / JPEG_CHUNK;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (it.part == 0) {
        for (uint32_t k = threadIdx.x; k < J.hdr_len; k += blockDim.x) J.out[k] = J.hdr[k];
        if (threadIdx.x == 0) { // EOI
            const uint32_t len = J.result[0];
            J.out[len - 2] = 0xff;
            J.out[len - 1] = 0xd9;
        }
    }
    for (uint32_t c = it.part * (JPEG_STUFF_THREADS / 32) + warp; c < n_chunks; c += it.nparts * (JPEG_STUFF_THREADS / 32)) {
        const uint4 v = *reinterpret_cast<const uint4 *>(J.words + (size_t)c * (JPEG_CHUNK / 4) + lane * 4);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
        const uint32_t n = ff_bytes(v.x) + ff_bytes(v.y) + ff_bytes(v.z) + ff_bytes(v.w);
        const uint32_t before = warp_inclusive(n, lane) - n;
        const uint32_t k0 = c * JPEG_CHUNK + lane * 16; // first scan byte of this lane
        uint8_t *o = J.out + J.hdr_len + k0 + J.chunk_off[c] + before;
#pragma unroll
        for (int k = 0; k < 16; k++) {
            if (k0 + k < U) {
                const uint32_t b = (w[k >> 2] >> (24 - 8 * (k & 3))) & 0xff;
                *o++ = (uint8_t)b;
                if (b == 0xff) *o++ = 0;
            }
        }
    }
}

} // namespace

cudaError_t launch_jpeg(const JpegJob *jobs, int n_jobs, const JpegDctItem *dct_items, int n_dct, const JpegStuffItem *stuff_items,
                        int n_stuff, cudaStream_t st)
{
    if (n_jobs <= 0) return cudaSuccess;
    k_jpeg_dct<<<n_dct, DCT_THREADS, 0, st>>>(jobs, dct_items);
    k_jpeg_offsets<<<n_jobs, JPEG_SCAN_THREADS, 0, st>>>(jobs);
    k_jpeg_zero<<<n_stuff, JPEG_STUFF_THREADS, 0, st>>>(jobs, stuff_items);
    k_jpeg_emit<<<n_dct, 192, 0, st>>>(jobs, dct_items);
    k_jpeg_ffcount<<<n_stuff, JPEG_STUFF_THREADS, 0, st>>>(jobs, stuff_items);
    k_jpeg_chunks<<<n_jobs, JPEG_SCAN_THREADS, 0, st>>>(jobs);
    k_jpeg_write<<<n_stuff, JPEG_STUFF_THREADS, 0, st>>>(jobs, stuff_items);
    return cudaGetLastError();
}

} // namespace ipg
