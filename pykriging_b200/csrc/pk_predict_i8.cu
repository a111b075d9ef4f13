// |L^-1 psi_q|^2 on the INT8 tensor cores: the predictor's dense contraction (predicterr_normalized, matrixops.py:79-95:
// psi^T Psi^-1 psi = |L^-1 psi|^2; expimp, krige.py:201-234) as an FP64-exact-product GEMM through tcgen05.mma.kind::i8
// with accumulators in TMEM.
//
// There is no FP64 path on the 5th-generation tensor cores (DMMA runs on the legacy pipe: 37 TFLOP/s on B200), but their
// INT8 path does 4.5 POPS with exact INT32 accumulation.  Ozaki splitting turns an FP64 product into integer ones: every
// operand is written as a sum of S = 7 balanced base-256 digits of a row-scaled fixed-point number,
//     x = 2^e (d_1 2^-8 + d_2 2^-16 + ... + d_7 2^-56),   -128 <= d_t <= 127,
// which is EXACT for the 56 bits below 2^e (a row of L^-1 takes the exponent of its largest entry; psi lies in [0, 1] and
// simply contributes the 7 bytes of round(psi 2^54) as UNSIGNED digits -- the psi kernels write them instead of doubles).  The digit-pair products with t + u <= 8 (28 of 49; the rest lie
// below 2^-58 of the row scales) are INT8 MMAs whose INT32 sums are exact and order independent; pairs of equal weight
// share one TMEM accumulator (7 levels x 64 columns = 448 of the 512 TMEM columns), and the epilogue adds the levels in
// FP64, squares and accumulates.  (Base 256 instead of 128: 7 digits and 28 products instead of 8 and 36 -- the kernel is
// bound by the tensor cores' shared-memory operand reads, 6 KB per MMA, so every product saved is time saved.)  Result: the accuracy of an FP64 dot product with NO rounding inside the sum (tools/ozaki_probe.cu: max
// error 1.9e-15 on sums of magnitude 32 at K = 512, below FP64's own eps sqrt K) at 73 FP64-equivalent TFLOP/s -- twice
// the DMMA peak.  A query's result depends on its own psi row only, so tiles, blocks and shards stay bit-identical.
//
// Pipeline of one 128-query tile (one CTA per SM, 192 threads): for every block of 64 rows of L^-1 (= 64 entries of
// v = L^-1 psi), K runs over the samples up to the diagonal in chunks of 64; a chunk is 8 digit planes of psi (128 x 64
// bytes each) and 8 of L^-1 (64 x 64 bytes), loaded by TMA with 64-byte swizzle into 2 stages of 96 KB; ONE thread issues
// the 72 MMAs of a chunk (128 x 64 x 32 each) and commits them to the stage's mbarrier; four epilogue warps (one TMEM
// lane = one query per thread) read the 8 levels, rebuild v in FP64 and add v^2.  After the last block: s and EI.
// (Eight epilogue warps since the accumulators cannot be double buffered -- they fill the TMEM -- and the MMAs of the
// next block wait for the epilogue: two warps share a lane quarter, each takes half of the columns.)
#include <cuda.h>

#include "pk_i8.cuh"
#include "pk_internal.h"

namespace pk {
namespace i8 {

constexpr int BM = 128;              // queries per tile (TMEM lanes)
constexpr int BN = 64;               // rows of L^-1 per block (TMEM columns per level)
constexpr int BK = 64;               // samples (bytes) per pipeline chunk: one 64-byte swizzle row
constexpr int STAGES = 2;
constexpr int A_SLICE = BM * BK, B_SLICE = BN * BK;
constexpr int STAGE_BYTES = S * (A_SLICE + B_SLICE);   // 96 KB
constexpr int THREADS = 320;         // warps 0-3 and 6-9: epilogue (a warp reads the TMEM lanes 32 (w % 4) .. + 31: two warps per
                                     // lane quarter, 32 of the 64 columns each); warp 4: TMA + TMEM allocation; warp 5: MMA
constexpr int EPI_THREADS = 256;

// L^-1 (np x np, row-major, lower) -> planes[t][i][k] and the row scales 2^e_i / PSI_SCALE: one warp per row
__global__ void __launch_bounds__(256)
linv_slice_kernel(const double* __restrict__ linv, int np, int8_t* __restrict__ planes, double* __restrict__ rowscale) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= np) return;
    const double* src = linv + (size_t)row * np;
    double amax = 0.0;
    for (int k = lane; k < np; k += 32) amax = fmax(amax, fabs(src[k]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) amax = fmax(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    // amax 2^-e in [0.125, 0.25): every entry of the row scaled by 2^-e is below 0.25 in magnitude
    const int e = (amax > 0.0 && amax < CUDART_INF) ? ilogb(amax) + 3 : 0;
    const double down = ldexp(1.0, -e);
    for (int k = lane; k < np; k += 32) {
        int d[S];
        const double x = src[k] * down;  // exact (power of two)
        balanced_digits((fabs(x) <= 0.25) ? x : 0.0, d);  // NaN / inf rows (a failed fit never gets here) -> zeros
#pragma unroll
        for (int t = 0; t < S; ++t) planes[((size_t)t * np + row) * np + k] = (int8_t)d[t];
    }
    if (lane == 0) rowscale[row] = ldexp(1.0, e) / PSI_SCALE;
}

__global__ void __launch_bounds__(THREADS, 1)
trinorm_i8_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int np, int64_t mc,
                  const double* __restrict__ rowscale, double sigma2, double y_min, double ei_weight, double var_extra,
                  const double* __restrict__ yhat, double* __restrict__ s_out, double* __restrict__ ei_out) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024 - ((unsigned)__cvta_generic_to_shared(smem_raw) & 1023)) & 1023);
    __shared__ unsigned long long full[STAGES], empty[STAGES], tmem_full, tmem_empty;
    __shared__ uint32_t tmem_base_s;
    __shared__ double scale_s[2][BN];
    __shared__ double qhalf_s[BM];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full + s, 1);
            mbar_init(empty + s, 1);
        }
        mbar_init(&tmem_full, 1);
        mbar_init(&tmem_empty, EPI_THREADS);
        mbar_fence_init();
    }
    if (warp == 4) {
        // 7 levels x 64 columns = 448; allocations are powers of two
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;\n" ::"r"(
            (unsigned)__cvta_generic_to_shared(&tmem_base_s)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    const int nb = np / BN;                                   // blocks of 64 rows of L^-1
    const int64_t ntiles = (mc + BM - 1) / BM;
    if (warp == 4 && lane == 0) {
        // ---- TMA producer: chunk kb of block c of tile -> stage (running chunk counter) ----
        unsigned it = 0;
        for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const int q0 = (int)(tile * BM);
            for (int c = 0; c < nb; ++c) {
                for (int kb = 0; kb <= c; ++kb, ++it) {
                    const int st = (int)(it % STAGES);
                    if (it >= STAGES) mbar_wait(empty + st, ((it / STAGES) - 1) & 1);
                    unsigned char* sa = smem + st * STAGE_BYTES;
                    unsigned char* sb = sa + S * A_SLICE;
                    mbar_arrive_expect_tx(full + st, (unsigned)STAGE_BYTES);
#pragma unroll
                    for (int t = 0; t < S; ++t) tma_load_3d(sa + t * A_SLICE, &tmA, kb * BK, q0, t, full + st);
#pragma unroll
                    for (int t = 0; t < S; ++t) tma_load_3d(sb + t * B_SLICE, &tmB, kb * BK, c * BN, t, full + st);
                }
            }
        }
    } else if (warp == 5 && lane == 0) {
        // ---- MMA issuer ----
        // A (psi digits) unsigned, B (L^-1 digits) signed
        const uint32_t idesc = idesc_i8(false, true, BM, BN);
        unsigned it = 0, ch = 0;
        for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            for (int c = 0; c < nb; ++c, ++ch) {
                if (ch > 0) {  // the epilogue has read the previous block's accumulators
                    mbar_wait(&tmem_empty, (ch - 1) & 1);
                    tc_fence_after();
                }
                for (int kb = 0; kb <= c; ++kb, ++it) {
                    const int st = (int)(it % STAGES);
                    mbar_wait(full + st, (it / STAGES) & 1);
                    tc_fence_after();
                    const uint32_t sa = (unsigned)__cvta_generic_to_shared(smem + st * STAGE_BYTES);
                    const uint32_t sb = sa + S * A_SLICE;
#pragma unroll
                    for (int kk = 0; kk < BK / 32; ++kk) {
#pragma unroll
                        for (int t = 0; t < S; ++t) {
#pragma unroll
                            for (int u = 0; u + t < S; ++u) {
                                // digits t+1 of psi and u+1 of L^-1, weight 2^-8(t+u+2): level t + u.  In the first chunk of a
                                // block the pairs with t = 0 touch every level first and overwrite it.
                                umma_i8(tmem_base + (t + u) * BN, smem_desc_k_sw64(sa + t * A_SLICE + kk * 32),
                                        smem_desc_k_sw64(sb + u * B_SLICE + kk * 32), idesc, (kb == 0 && kk == 0 && t == 0) ? 0u : 1u);
                            }
                        }
                    }
                    umma_commit(empty + st);  // the stage is free once these MMAs have read it
                }
                umma_commit(&tmem_full);      // ... and the block's accumulators are complete
            }
        }
    } else if (warp < 4 || warp >= 6) {
        // ---- epilogue: one query per thread pair; the thread of warps 0-3 takes columns 0-31 of a block, its partner in
        // warps 6-9 columns 32-63 ----
        const int quarter = warp & 3;                 // the TMEM lanes this warp may read
        const int colhalf = (warp < 4) ? 0 : 1;
        const int etid = (warp < 4) ? tid : tid - 64; // 0 .. 255 over the epilogue threads
        const int row = quarter * 32 + lane;          // query within the tile
        unsigned ch = 0;
        for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const int64_t q = tile * BM + row;
            double qs = 0.0;
            for (int c = 0; c < nb; ++c, ++ch) {
                // the block's row scales, double buffered in shared memory
                if (etid < BN) scale_s[ch & 1][etid] = rowscale[c * BN + etid];
                asm volatile("bar.sync 1, 256;\n" ::: "memory");
                mbar_wait(&tmem_full, ch & 1);
                tc_fence_after();
                const uint32_t lane_base = tmem_base + ((uint32_t)(quarter * 32) << 16) + colhalf * 32;
#pragma unroll 1
                for (int part = 0; part < 2; ++part) {
                    double v[16];
                    levels_to_fp64(lane_base, part * 16, BN, v);
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const double x = v[j] * scale_s[ch & 1][colhalf * 32 + part * 16 + j];
                        qs = fma(x, x, qs);
                    }
                }
                tc_fence_before();
                mbar_arrive(&tmem_empty);
            }
            // the two halves of a query's sum meet in shared memory (fixed order: columns 0-31 first)
            if (colhalf == 1) qhalf_s[row] = qs;
            asm volatile("bar.sync 2, 256;\n" ::: "memory");
            if (colhalf == 0 && q < mc) {
                qs += qhalf_s[row];
                if (yhat[q] != yhat[q]) qs = yhat[q];  // a NaN query has no digits: let the NaN of its yhat through
                const double ssqr = sigma2 * (1.0 + var_extra - qs);
                const double Sd = sqrt(fabs(ssqr));
                if (s_out) s_out[q] = Sd;
                if (ei_out) ei_out[q] = expected_improvement(yhat[q], Sd, y_min, ei_weight);
            }
            asm volatile("bar.sync 2, 256;\n" ::: "memory");  // qhalf_s may be overwritten by the next tile
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(tmem_base));
}

static cudaError_t encode_planes(CUtensorMap* tm, const void* base, uint64_t cols, uint64_t rows, uint64_t plane_stride,
                                 uint32_t box_rows) {
    typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static encode_fn encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
        if (e != cudaSuccess) return e;
        if (qres != cudaDriverEntryPointSuccess || !fn) return cudaErrorNotSupported;
        encode = (encode_fn)fn;
    }
    const cuuint64_t gdim[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)S};
    const cuuint64_t gstride[2] = {(cuuint64_t)cols, (cuuint64_t)plane_stride};
    const cuuint32_t box[3] = {(cuuint32_t)BK, box_rows, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = encode(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void*>(base), gdim, gstride, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return (r == CUDA_SUCCESS) ? cudaSuccess : cudaErrorInvalidValue;
}

}  // namespace i8

// Digit planes + row scales of the cached L^-1 (once per fit / append).
cudaError_t ensure_linv_i8(pk_handle* h) {
    const int np = h->fit_np;
    const size_t need = (size_t)i8::S * np * np;
    cudaError_t e;
    if (need > h->cap_linv_i8) {
        e = cudaStreamSynchronize(h->stream);
        if (e != cudaSuccess) return e;
        if (h->d_linv_i8) cudaFree(h->d_linv_i8);
        if (h->d_linv_scale) cudaFree(h->d_linv_scale);
        h->d_linv_i8 = nullptr;
        h->d_linv_scale = nullptr;
        h->cap_linv_i8 = 0;
        e = cudaMalloc((void**)&h->d_linv_i8, need);
        if (e != cudaSuccess) return e;
        e = cudaMalloc((void**)&h->d_linv_scale, sizeof(double) * np);
        if (e != cudaSuccess) return e;
        h->cap_linv_i8 = need;
        h->linv_i8_valid = false;
    }
    if (!h->linv_i8_valid) {
        i8::linv_slice_kernel<<<(np + 7) / 8, 256, 0, h->stream>>>(h->d_linv, np, h->d_linv_i8, h->d_linv_scale);
        h->launches++;
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        h->linv_i8_valid = true;
    }
    return cudaSuccess;
}

// Digit planes of psi for a block of up to `mcap` queries: (plane pointer, bytes between planes).  The psi kernels of the
// split pipeline write them directly.
cudaError_t psi_planes_i8(pk_handle* h, int64_t mcap, uint8_t** planes, size_t* plane_stride) {
    const size_t stride = (size_t)mcap * h->fit_np;
    const size_t need = (size_t)i8::S * stride;
    if (need > h->cap_psi_i8) {
        cudaError_t e = cudaStreamSynchronize(h->stream);
        if (e != cudaSuccess) return e;
        if (h->d_psi_i8) cudaFree(h->d_psi_i8);
        h->d_psi_i8 = nullptr;
        h->cap_psi_i8 = 0;
        e = cudaMalloc((void**)&h->d_psi_i8, need);
        if (e != cudaSuccess) return e;
        h->cap_psi_i8 = need;
    }
    *planes = reinterpret_cast<uint8_t*>(h->d_psi_i8);
    *plane_stride = stride;
    return cudaSuccess;
}

// s / EI of `mc` queries whose psi digit planes are in place (psi_planes_i8): the INT8 contraction.
cudaError_t launch_trinorm_i8(pk_handle* h, int64_t mc, int64_t mcap, double sigma2, double y_min, double ei_weight,
                              double var_extra, const double* yhat, double* sout, double* ei) {
    const int np = h->fit_np;
    cudaError_t e = ensure_linv_i8(h);
    if (e != cudaSuccess) return e;
    const size_t plane_stride = (size_t)mcap * np;
    const size_t smem = (size_t)i8::STAGES * i8::STAGE_BYTES + 1024;
    static bool attr_done_dev[PK_MAX_DEVICES] = {};
    bool& attr_done = attr_done_dev[h->device & (PK_MAX_DEVICES - 1)];
    if (!attr_done) {
        e = cudaFuncSetAttribute(i8::trinorm_i8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attr_done = true;
    }
    CUtensorMap tmA, tmB;
    e = i8::encode_planes(&tmA, h->d_psi_i8, (uint64_t)np, (uint64_t)mc, (uint64_t)plane_stride, i8::BM);
    if (e != cudaSuccess) return e;
    e = i8::encode_planes(&tmB, h->d_linv_i8, (uint64_t)np, (uint64_t)np, (uint64_t)np * np, i8::BN);
    if (e != cudaSuccess) return e;
    const int64_t ntiles = (mc + i8::BM - 1) / i8::BM;
    const int grid = (int)((ntiles < (int64_t)h->sm_count) ? ntiles : (int64_t)h->sm_count);
    i8::trinorm_i8_kernel<<<grid, i8::THREADS, smem, h->stream>>>(tmA, tmB, np, mc, h->d_linv_scale, sigma2, y_min, ei_weight,
                                                                 var_extra, yhat, sout, ei);
    h->launches++;
    return cudaGetLastError();
}

}  // namespace pk
