// Implicit-GEMM convolution, forward, for sm_100a: TMA-staged NHWC tiles -> tcgen05.mma (TMEM accumulators)
// -> fused epilogue.  One persistent CTA per SM, warp-specialised:
//   warp 0   TMA producer   (one elected lane)
//   warp 1   MMA issuer     (one elected lane; warp 3 is a second issuer on small-N launches)
//   warp 2   TMEM allocator
//   warps 4-11 epilogue     (TMEM -> registers -> alpha/bias/activation/mul -> global)
//
// GEMM view (reference call sites: models.py:83,94,162,174; torchvision vgg19 convs, losses.py:43):
//   M = 128 output pixels (a tn x th x tw box of the output), N = block_n output channels,
//   K = for each filter row r: the kw*C contiguous input values (s, c) of the padded NHWC row, cut into
//       128-byte chunks (32 tf32 / 64 bf16 values).  The A tile of one chunk is ONE 5-D TMA box
//       {chunk, tw, 1, th, tn} over the tensor map {window, wo, r, ho, n} with byte strides
//       {es, stride*C*es, row, stride*row, img}: sliding windows are expressed by overlapping strides, so any
//       kernel size / stride 1-2 / channel count (C*es % 16 == 0) is the same code path, and the landing
//       layout (128 rows x 128 B, SWIZZLE_128B) is exactly the canonical K-major UMMA operand.
//   Padding is never materialised by this kernel: the producer of x wrote the halo (reflect or zero).
#include <type_traits>

#include "common.cuh"
#include "host_util.h"

namespace uegan {

constexpr int kMaxStages = 8;
constexpr int kMaxBSlots = 16;
constexpr int kABytes = 128 * 128;  // 128 pixel rows x 128 B
constexpr int kTmemCols = 512;      // 2 accumulator stages x 256 fp32 columns

struct ConvParams {
  // tiling
  int tw, th, tn;
  int tw_log2, th_log2;
  int tiles_w, tiles_h, tiles_img;  // number of M tiles along wo / ho / n
  int n_tiles, block_n;
  int total_tiles;
  // K loop
  int kh, chunks_per_row, chunk_elems, num_k_chunks;
  int num_stages, stage_bytes;
  // patch mode (stride-1, pixel = whole 128-byte chunks, weights resident in smem)
  int patch_w, patch_nch, patch_k, patch_off;  // patch width in pixels, chunks per pixel, kernel size, halo - pad
  // patch-mode operand rows: 128 bytes (SWIZZLE_128B, 4 MMAs of K = 32 bytes per tap) or, for 32-channel 16-bit tensors,
  // 64 bytes = one pixel (SWIZZLE_64B, 2 MMAs per tap); weight-tile K offset = (r * patch_cpr + s * patch_nch + c) * patch_ce
  int patch_row_bytes, patch_kmma, patch_layout, patch_cpr, patch_ce;
  int n_issuers;  // 1, or 2: warps 1 and 3 both issue MMAs, alternating tiles (small-N launches)
  int w_tile_bytes, w_total_bytes, stage_tx_bytes;
  // patch mode with STREAMED weights (kPatch == 2): the weight tiles do not fit next to the patches (large C or N), so
  // they flow through their own ring of b_slots x w_tile_bytes behind the A ring (b_ring_off bytes from the smem base)
  int b_slots, b_ring_off;
  // epilogue
  int Wo, Ho, Nimg, cout;
  int act;
  int out_kind;  // storage of y / mul: UEGAN_F32 (tf32-rounded), UEGAN_BF16, UEGAN_F16
  int ab_fmt;    // UMMA operand format of x and w
  void* out;  // points at element (n=0, y=0, x=0, c=y_c_off) of the interior
  long long out_pix, out_row, out_img;  // strides in elements
  const float* bias;
  const float* alpha;
  // per-tensor power-of-two scales (uegan_tensor.scale; NULL = 1): stored = scale * true
  const float* sx;    // input activations
  const float* sw;    // packed weights
  const float* sy;    // NHWC output (planar fp32 outputs are true values)
  const float* smul;  // the `mul` operand
  const void* mul;  // same dtype as out
  long long mul_pix, mul_row, mul_img;
  // optional second NHWC output: the values BEFORE `mul` (training keeps y4 next to y4 * x1, models.py:70), own scale
  void* premul;
  long long pre_pix, pre_row, pre_img;
  const float* spre;
  const void* mask;  // activation-derivative mask from a forward tensor (dgrad epilogue); dtype mask_kind
  long long mask_pix, mask_row, mask_img;
  int mask_kind, mask_act;
  float* out_nchw;
  const float* residual_nchw;
  float* aux_nchw;  // optional: activation before the residual/clamp
  unsigned int* err_sink;  // host-mapped watchdog word
  double* in_stats;        // optional [Nimg][cout][2] sum / sum-of-squares of the stored outputs (InstanceNorm)
  // merged parity classes of a stride-2 data gradient (uegan_conv_desc.y_cls_c): column = (class, channel), class
  // (pi, pj) lands at output pixel (2a + pi, 2b + pj); cls_row / cls_pix = element strides of ONE output row / pixel
  int cls_c, cls_h, cls_w;
  long long cls_row, cls_pix;
  // reflect_halo = p > 0: the epilogue also writes the reflection-padding halo of y (nn.ReflectionPad2d of the CONSUMER,
  // models.py:82,93,161,173): the thread that owns interior pixel (i, j) stores its value at every halo position that
  // mirrors it (rows -i / 2(H-1)-i for i within p of an edge, same for columns) -- no separate halo pass over the tensor
  int reflect_halo;
};

__device__ __forceinline__ float apply_act(float v, int act) {
  switch (act) {
    case UEGAN_ACT_LRELU: return v > 0.f ? v : 0.2f * v;
    case UEGAN_ACT_RELU: return fmaxf(v, 0.f);
    case UEGAN_ACT_TANH: return tanhf(v);
    case UEGAN_ACT_SIGMOID: return 1.f / (1.f + __expf(-v));
    default: return v;
  }
}
__device__ __forceinline__ float round_tf32(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return __uint_as_float(r);
}

// value as it will be stored (so statistics and stores agree)
__device__ __forceinline__ float round_store(float v, int kind) {
  if (kind == UEGAN_F32) return round_tf32(v);
  if (kind == UEGAN_BF16) return __bfloat162float(__float2bfloat16_rn(v));
  return __half2float(__float2half_rn(v));
}

struct TileCoord {
  int wo0, ho0, n0, nt;
};
__device__ __forceinline__ TileCoord decode_tile(const ConvParams& p, int t) {
  TileCoord c;
  c.nt = t % p.n_tiles;
  int m = t / p.n_tiles;
  c.wo0 = (m % p.tiles_w) * p.tw;
  m /= p.tiles_w;
  c.ho0 = (m % p.tiles_h) * p.th;
  c.n0 = (m / p.tiles_h) * p.tn;
  return c;
}

// The tiles a persistent CTA visits (t = blockIdx.x, += gridDim.x) without a division per tile: the stride is decomposed
// once into its (n-tile, w, h, image) digits and added with carries.  The epilogue warps are bound by their own
// instruction latency on small-N layers (r2r: ~340 warp instructions per 128 x 32 tile, 90 of them three integer
// divisions), so every instruction per tile counts there.
struct TileIter {
  int i_nt, i_w, i_h, i_n;  // current digits
  int s_nt, s_w, s_h, s_n;  // digits of the stride
  __device__ __forceinline__ void init(const ConvParams& p, int t0, int step) {
    i_nt = t0 % p.n_tiles; int m = t0 / p.n_tiles;
    i_w = m % p.tiles_w; m /= p.tiles_w;
    i_h = m % p.tiles_h; i_n = m / p.tiles_h;
    s_nt = step % p.n_tiles; m = step / p.n_tiles;
    s_w = m % p.tiles_w; m /= p.tiles_w;
    s_h = m % p.tiles_h; s_n = m / p.tiles_h;
  }
  __device__ __forceinline__ void next(const ConvParams& p) {
    i_nt += s_nt; int c = i_nt >= p.n_tiles; i_nt -= c ? p.n_tiles : 0;
    i_w += s_w + c; c = i_w >= p.tiles_w; i_w -= c ? p.tiles_w : 0;
    i_h += s_h + c; c = i_h >= p.tiles_h; i_h -= c ? p.tiles_h : 0;
    i_n += s_n + c;
  }
  __device__ __forceinline__ TileCoord coord(const ConvParams& p) const {
    TileCoord c;
    c.nt = i_nt; c.wo0 = i_w * p.tw; c.ho0 = i_h * p.th; c.n0 = i_n * p.tn;
    return c;
  }
};

template <int ACT>
__device__ __forceinline__ float act_t(float v, int act_rt) {
  if constexpr (ACT == UEGAN_ACT_LRELU) return fmaxf(v, 0.2f * v);
  else if constexpr (ACT == UEGAN_ACT_RELU) return fmaxf(v, 0.f);
  else if constexpr (ACT == UEGAN_ACT_NONE) return v;
  else return apply_act(v, act_rt);
}

// NHWC epilogue of one warp: its 32 rows x the 16-column chunks {half, half+2, ...} of the tile.
template <int ACT, int NH>
__device__ __forceinline__ void epilogue_nhwc(const ConvParams& p, uint32_t taddr, int colbase, int half, long long o,
                                              long long mo, bool valid, float alpha, float bmul, float* stat_slice,
                                              int mk_n, int mk_h, int mk_w, const float* s_bias, long long po,
                                              float pre_scale) {
  for (int c0 = half * 16; c0 < p.block_n; c0 += 16 * NH) {
    uint32_t rr[16];
    tmem_ld16(taddr + c0, rr);
    tmem_ld_wait();
    const int col0 = colbase + c0;
    if (col0 >= p.cout) continue;  // warp-uniform
    long long oc = o + c0;
    bool vc = valid;
    if (p.cls_c) {
      const int cls = col0 / p.cls_c, ch = col0 - cls * p.cls_c, pi = cls >> 1, pj = cls & 1;
      oc = o - colbase + pi * p.cls_row + pj * p.cls_pix + ch;
      vc = valid && (2 * mk_h + pi < p.cls_h) && (2 * mk_w + pj < p.cls_w);
    }
    if (p.in_stats) {
      // per-(n, c) sum and sum of squares over this warp's 32 rows (all rows of a tile share n when tn == 1):
      // recursive-halving butterfly, 31 shuffles for 32 values; lane L ends with the total of element L.
      float a[32];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        float x = __uint_as_float(rr[i]) * alpha;
        if (p.bias) x += __ldg(p.bias + col0 + i) * bmul;
        x = act_t<ACT>(x, p.act);
        x = round_store(x, p.out_kind);
        x = valid ? x : 0.f;
        a[i] = x;
        a[16 + i] = x * x;
      }
#pragma unroll
      for (int ofs = 16, len = 16; ofs >= 1; ofs >>= 1, len >>= 1) {
        const bool up = (threadIdx.x & ofs) != 0;
#pragma unroll
        for (int i = 0; i < len; ++i) {
          const float send = up ? a[i] : a[i + len];
          const float keep = up ? a[i + len] : a[i];
          a[i] = keep + __shfl_xor_sync(0xffffffffu, send, ofs);
        }
      }
      stat_slice[(c0 >> 5) * 32 + (threadIdx.x & 31)] += a[0];  // private to this lane; flushed when n changes
    }
    if (!vc) continue;
    float v[16];
    if (p.bias) {
      // bias * bmul staged in shared memory once per CTA (bmul = 1 unless the output carries a scale: a power of two, so
      // the product is exact); a global load per tile sat on the epilogue's critical path
      const float4* bp = reinterpret_cast<const float4*>(s_bias + col0);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 b = bp[i];
        v[4 * i + 0] = act_t<ACT>(fmaf(__uint_as_float(rr[4 * i + 0]), alpha, b.x), p.act);
        v[4 * i + 1] = act_t<ACT>(fmaf(__uint_as_float(rr[4 * i + 1]), alpha, b.y), p.act);
        v[4 * i + 2] = act_t<ACT>(fmaf(__uint_as_float(rr[4 * i + 2]), alpha, b.z), p.act);
        v[4 * i + 3] = act_t<ACT>(fmaf(__uint_as_float(rr[4 * i + 3]), alpha, b.w), p.act);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = act_t<ACT>(__uint_as_float(rr[i]) * alpha, p.act);
    }
    if (p.premul) {  // 16-bit outputs only (checked on the host)
      uint32_t pk[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (p.out_kind == UEGAN_BF16) {
          __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i] * pre_scale, v[2 * i + 1] * pre_scale);
          pk[i] = *reinterpret_cast<uint32_t*>(&h);
        } else {
          asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(pk[i]) : "f"(v[2 * i + 1] * pre_scale), "f"(v[2 * i] * pre_scale));
        }
      }
      uint4* op = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(p.premul) + po + c0);
      op[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      op[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
    }
    if (p.mul) {
      if (p.out_kind != UEGAN_F32) {
        const uint4* mp = reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(p.mul) + mo + c0);
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const uint4 qv = __ldg(mp + i);
          const uint32_t w[4] = {qv.x, qv.y, qv.z, qv.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (p.out_kind == UEGAN_BF16) {
              const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&w[j]);
              v[8 * i + 2 * j] *= __low2float(h);
              v[8 * i + 2 * j + 1] *= __high2float(h);
            } else {
              const __half2 h = *reinterpret_cast<const __half2*>(&w[j]);
              v[8 * i + 2 * j] *= __low2float(h);
              v[8 * i + 2 * j + 1] *= __high2float(h);
            }
          }
        }
      } else {
        const float4* mp = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.mul) + mo + c0);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 qv = __ldg(mp + i);
          v[4 * i + 0] *= qv.x; v[4 * i + 1] *= qv.y; v[4 * i + 2] *= qv.z; v[4 * i + 3] *= qv.w;
        }
      }
    }
    if (p.mask) {
      const long long ko = (long long)mk_n * p.mask_img + (long long)mk_h * p.mask_row + (long long)mk_w * p.mask_pix + colbase + c0;
      float t[16];
      if (p.mask_kind == UEGAN_F32) {
        const float4* mp = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.mask) + ko);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 qv = __ldg(mp + i);
          t[4 * i] = qv.x; t[4 * i + 1] = qv.y; t[4 * i + 2] = qv.z; t[4 * i + 3] = qv.w;
        }
      } else {  // 16-bit mask: only the sign matters, which is bit 15 of either 16-bit format
        const uint4* mp = reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(p.mask) + ko);
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const uint4 qv = __ldg(mp + i);
          const uint32_t w[4] = {qv.x, qv.y, qv.z, qv.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t lo = w[j] & 0xFFFFu, hi = w[j] >> 16;
            // value > 0  <=>  sign bit clear and not (+/-)zero
            t[8 * i + 2 * j] = (lo != 0u && lo < 0x8000u) ? 1.f : 0.f;
            t[8 * i + 2 * j + 1] = (hi != 0u && hi < 0x8000u) ? 1.f : 0.f;
          }
        }
      }
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        if (p.mask_act == UEGAN_ACT_LRELU) v[i] *= (t[i] > 0.f ? 1.f : 0.2f);
        else v[i] = t[i] > 0.f ? v[i] : 0.f;
      }
    }
    if (p.out_kind != UEGAN_F32) {
      uint32_t pk[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (p.out_kind == UEGAN_BF16) {
          __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
          pk[i] = *reinterpret_cast<uint32_t*>(&h);
        } else {  // saturating (see Vec<__half>::store): one F2FP.SATFINITE instead of four FMNMX + F2FP per pair
          asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(pk[i]) : "f"(v[2 * i + 1]), "f"(v[2 * i]));
        }
      }
      uint4* op = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(p.out) + oc);
      op[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      op[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
      if (p.reflect_halo) {
        // mirrors of (mk_h, mk_w): at most one per axis (host: H, W > 2 * halo + 1)
        const int hp = p.reflect_halo;
        const int dr = (mk_h >= 1 && mk_h <= hp) ? -2 * mk_h : ((mk_h <= p.Ho - 2 && mk_h >= p.Ho - 1 - hp) ? 2 * (p.Ho - 1 - mk_h) : 0);
        const int dc = (mk_w >= 1 && mk_w <= hp) ? -2 * mk_w : ((mk_w <= p.Wo - 2 && mk_w >= p.Wo - 1 - hp) ? 2 * (p.Wo - 1 - mk_w) : 0);
        if (dr | dc) {
          uint16_t* ob = reinterpret_cast<uint16_t*>(p.out) + oc;
          if (dr) {
            uint4* o2 = reinterpret_cast<uint4*>(ob + (long long)dr * p.out_row);
            o2[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]); o2[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
          }
          if (dc) {
            uint4* o2 = reinterpret_cast<uint4*>(ob + (long long)dc * p.out_pix);
            o2[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]); o2[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
          }
          if (dr && dc) {
            uint4* o2 = reinterpret_cast<uint4*>(ob + (long long)dr * p.out_row + (long long)dc * p.out_pix);
            o2[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]); o2[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
          }
        }
      }
    } else {
      float4* op = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + oc);
#pragma unroll
      for (int i = 0; i < 4; ++i)
        op[i] = make_float4(round_tf32(v[4 * i]), round_tf32(v[4 * i + 1]), round_tf32(v[4 * i + 2]),
                            round_tf32(v[4 * i + 3]));
    }
  }
}

// planar fp32 NCHW epilogue (cout <= 16): D prediction heads, G's last conv (tanh, + residual, clamp)
__device__ __forceinline__ void epilogue_planar(const ConvParams& p, uint32_t taddr, int colbase, int n, int ho, int wo,
                                                bool valid, float alpha) {
  uint32_t rr[16];
  tmem_ld16(taddr, rr);
  tmem_ld_wait();
  if (!valid) return;
  const long long plane = (long long)p.Ho * p.Wo;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    if (colbase + i < p.cout) {
      const long long o = ((long long)n * p.cout + colbase + i) * plane + (long long)ho * p.Wo + wo;
      float x = __uint_as_float(rr[i]) * alpha;
      if (p.bias) x += __ldg(p.bias + colbase + i);
      x = apply_act(x, p.act);
      if (p.aux_nchw) p.aux_nchw[o] = x;
      if (p.residual_nchw) x = fminf(fmaxf(x + __ldg(p.residual_nchw + o), -1.f), 1.f);
      p.out_nchw[o] = x;
    }
  }
}

// kOcc2: the two-CTAs-per-SM build for small-N launches (block_n <= 128): 256 threads (4 epilogue warps instead of 8),
// <= 128 registers, 256 TMEM columns (2 accumulator stages x 128), <= ~100 KB of shared memory.  Those launches are bound
// by the latencies of their one-thread producer / issuer and of the per-tile epilogue chain, not by any throughput; a
// second resident CTA is a second, independent tile pipeline on the same SM.
template <int kTf32, int kPatch, int kOcc2>
__global__ void __launch_bounds__(kOcc2 ? 256 : 384, kOcc2 ? 2 : 1)
conv_fprop_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const ConvParams p) {
  constexpr int kAccCols = kOcc2 ? 128 : 256;      // TMEM columns per accumulator stage
  constexpr int kEpiWarps = kOcc2 ? 4 : 8;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t full_bar[kMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kMaxStages];
  __shared__ __align__(8) uint64_t tmem_full[2];
  __shared__ __align__(8) uint64_t tmem_empty[2];
  __shared__ __align__(8) uint64_t w_full;
  __shared__ __align__(8) uint64_t fullB[kMaxBSlots];
  __shared__ __align__(8) uint64_t emptyB[kMaxBSlots];
  __shared__ uint32_t tmem_base_smem;
  __shared__ float s_stats[8 * 256];  // InstanceNorm partial sums: [epilogue warp][chunk slot][lane]
  __shared__ __align__(16) float s_bias[512];  // bias * (s_y / s_mul), padded with zeros to the 16-column chunks

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && elect_one()) {
    for (int i = 0; i < p.num_stages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], kEpiWarps);
    }
    mbar_init(&w_full, 1);
    if constexpr (kPatch == 2) {
      for (int i = 0; i < p.b_slots; ++i) {
        mbar_init(&fullB[i], 1);
        mbar_init(&emptyB[i], 1);
      }
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(&tmem_base_smem, 2 * kAccCols);
  pdl_sync();  // everything above is on-chip set-up: it overlaps the tail of the preceding kernel
  if (p.bias && !p.out_nchw) {
    const float bm = (p.sy ? __ldg(p.sy) : 1.0f) / (p.smul ? __ldg(p.smul) : 1.0f);
    for (int i = threadIdx.x; i < 512; i += blockDim.x) s_bias[i] = i < p.cout ? __ldg(p.bias + i) * bm : 0.f;
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == 0) {
    if (elect_one()) {
      // ===================== TMA producer =====================
      int stage = 0;
      uint32_t phase = 0;
      if constexpr (kPatch == 2) {
        // A: one patch per (tile, channel chunk); B: the k*k weight tiles of that chunk, each through the B ring.  The MMA
        // warp consumes in exactly this order, so neither ring can starve the other.
        uint8_t* sb0 = smem + p.b_ring_off;
        int bs = 0;
        uint32_t bphase = 0;
        const int taps = p.patch_k * p.patch_k;
        for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
          const TileCoord tc = decode_tile(p, t);
          for (int c = 0; c < p.patch_nch; ++c) {
            mbar_wait(&empty_bar[stage], phase ^ 1, 0x100 + stage, p.err_sink);
            mbar_arrive_expect_tx(&full_bar[stage], (uint32_t)p.stage_tx_bytes);
            tma_load_4d(&tmA, &full_bar[stage], smem + stage * p.stage_bytes, c * p.chunk_elems, tc.wo0 + p.patch_off,
                        tc.ho0 + p.patch_off, tc.n0);
            if (++stage == p.num_stages) { stage = 0; phase ^= 1; }
            int r = 0, s_ = 0;
            for (int tap = 0; tap < taps; ++tap) {
              mbar_wait(&emptyB[bs], bphase ^ 1, 0x180 + bs, p.err_sink);
              mbar_arrive_expect_tx(&fullB[bs], (uint32_t)p.w_tile_bytes);
              const int kofs = (r * p.chunks_per_row + s_ * p.patch_nch + c) * p.chunk_elems;
              tma_load_2d(&tmB, &fullB[bs], sb0 + bs * p.w_tile_bytes, kofs, tc.nt * p.block_n);
              if (++s_ == p.patch_k) { s_ = 0; ++r; }
              if (++bs == p.b_slots) { bs = 0; bphase ^= 1; }
            }
          }
        }
      } else if constexpr (kPatch == 1) {
        // weights: all k*k*nch [block_n x 128 B] tiles once per CTA, resident behind the A ring
        uint8_t* sw = smem;
        mbar_arrive_expect_tx(&w_full, (uint32_t)p.w_total_bytes);
        // 16-byte pixels: one box per filter row, landing as [K chunk of 8][cout][8 values] = unswizzled core matrices
        const int ntiles_w = p.patch_layout == UMMA_LAYOUT_NONE ? 0 : p.patch_k * p.patch_k * p.patch_nch;
        if (p.patch_layout == UMMA_LAYOUT_NONE)
          for (int r = 0; r < p.patch_k; ++r) tma_load_4d(&tmB, &w_full, sw + r * p.w_tile_bytes, 0, 0, 0, r);
        for (int wi = 0; wi < ntiles_w; ++wi) {
          const int c = wi % p.patch_nch, tap = wi / p.patch_nch;
          const int r = tap / p.patch_k, s_ = tap % p.patch_k;
          const int kofs = (r * p.patch_cpr + s_ * p.patch_nch + c) * p.patch_ce;
          tma_load_2d(&tmB, &w_full, sw + wi * p.w_tile_bytes, kofs, 0);
        }
        uint8_t* sa0 = smem + p.w_total_bytes;
        for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
          const TileCoord tc = decode_tile(p, t);
          for (int c = 0; c < p.patch_nch; ++c) {
            mbar_wait(&empty_bar[stage], phase ^ 1, 0x100 + stage, p.err_sink);
            mbar_arrive_expect_tx(&full_bar[stage], (uint32_t)p.stage_tx_bytes);
            tma_load_4d(&tmA, &full_bar[stage], sa0 + stage * p.stage_bytes, c * p.patch_ce, tc.wo0 + p.patch_off,
                        tc.ho0 + p.patch_off, tc.n0);
            if (++stage == p.num_stages) { stage = 0; phase ^= 1; }
          }
        }
      } else {
        const uint32_t tx_bytes = kABytes + p.block_n * 128;
        for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
          const TileCoord tc = decode_tile(p, t);
          int r = 0, j = 0;
          for (int kc = 0; kc < p.num_k_chunks; ++kc) {
            mbar_wait(&empty_bar[stage], phase ^ 1, 0x100 + stage, p.err_sink);
            uint8_t* sa = smem + stage * p.stage_bytes;
            uint8_t* sb = sa + kABytes;
            mbar_arrive_expect_tx(&full_bar[stage], tx_bytes);
            tma_load_5d(&tmA, &full_bar[stage], sa, j * p.chunk_elems, tc.wo0, r, tc.ho0, tc.n0);
            tma_load_2d(&tmB, &full_bar[stage], sb, kc * p.chunk_elems, tc.nt * p.block_n);
            if (++j == p.chunks_per_row) { j = 0; ++r; }
            if (++stage == p.num_stages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1 || warp == 3) {
    // ===================== MMA issuers =====================
    // The issue loop is ONE thread's scalar instruction stream (descriptor arithmetic, R2UR moves into the uniform
    // registers tcgen05.mma reads): on small-N layers (N <= 64: 18 - 50 cheap MMAs per 128-pixel tile) that stream, not the
    // tensor pipe, was the launch's critical path (r2t: the issuing warp never waits; doing the same work on all 32 lanes
    // made the layers 30 % slower).  With n_iss = 2, warp 3 is a SECOND issuer: issuer i owns the CTA's tiles i, i + 2, ...
    // and the accumulator stage i; each walks the smem ring in tile order, skipping the other's stages.
    const int iss = warp == 3 ? 1 : 0;
    const int n_iss = (kPatch != 2 && p.n_issuers == 2) ? 2 : 1;
    if (iss < n_iss && elect_one()) {
      const uint32_t idesc = make_instr_desc(p.ab_fmt, 128, p.block_n);
      int stage = 0;
      uint32_t phase = 0;
      int acc = iss;
      uint32_t acc_phase = 0;
      const int per_tile = kPatch == 1 ? p.patch_nch : p.num_k_chunks;  // smem stages one tile consumes
      auto skip_tile = [&]() {
        for (int i = 0; i < per_tile; ++i)
          if (++stage == p.num_stages) { stage = 0; phase ^= 1; }
      };
      auto next_acc = [&]() {
        if (n_iss == 2) { acc_phase ^= 1; skip_tile(); }
        else { acc ^= 1; if (acc == 0) acc_phase ^= 1; }
      };
      if (iss) skip_tile();
      if constexpr (kPatch == 1) mbar_wait(&w_full, 0, 0x500, p.err_sink);
      int bs = 0;
      uint32_t bphase = 0;
      for (int t = blockIdx.x + iss * gridDim.x; t < p.total_tiles; t += n_iss * gridDim.x) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1, 0x200 + acc, p.err_sink);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + acc * kAccCols;
        if constexpr (kPatch == 2) {
          // same descriptor-window taps as the resident-weight patch mode; the B tile of every (chunk, tap) arrives
          // through the B ring and its slot is released by a commit right after the four MMAs that read it
          const uint32_t w_addr = smem_u32(smem);
          const uint32_t sbo = p.patch_w * 128;
          const uint64_t db0 = make_smem_desc(w_addr + p.b_ring_off, 16, 1024, UMMA_LAYOUT_SW128);
          uint32_t first = 0;
          for (int c = 0; c < p.patch_nch; ++c) {
            mbar_wait(&full_bar[stage], phase, 0x300 + stage, p.err_sink);
            tcgen05_fence_after();
            const uint64_t da0 = make_smem_desc(w_addr + stage * p.stage_bytes, 16, sbo, UMMA_LAYOUT_SW128);
            uint32_t a_off = 0;
            for (int r = 0; r < p.patch_k; ++r) {
              uint32_t a_rs = a_off;
              for (int s_ = 0; s_ < p.patch_k; ++s_) {
                mbar_wait(&fullB[bs], bphase, 0x380 + bs, p.err_sink);
                tcgen05_fence_after();
                const uint64_t da = desc_adv(da0, a_rs), db = desc_adv(db0, (uint32_t)bs * p.w_tile_bytes);
                umma_ss<kTf32>(d_tmem, da, db, idesc, first);
                umma_ss<kTf32>(d_tmem, da + 2, db + 2, idesc, 1u);
                umma_ss<kTf32>(d_tmem, da + 4, db + 4, idesc, 1u);
                umma_ss<kTf32>(d_tmem, da + 6, db + 6, idesc, 1u);
                umma_commit(&emptyB[bs]);
                if (++bs == p.b_slots) { bs = 0; bphase ^= 1; }
                first = 1;
                a_rs += 128;
              }
              a_off += sbo;
            }
            umma_commit(&empty_bar[stage]);
            if (++stage == p.num_stages) { stage = 0; phase ^= 1; }
          }
        } else if constexpr (kPatch == 1) {
          // one staged patch per 128-byte channel chunk; every filter tap (r, s) is the SAME smem patch read
          // through a descriptor window shifted by (r*patch_w + s) rows (tcgen05 swizzles on absolute address bits,
          // profiles/r1_probe_umma_window.json), 8-row groups = one output row of 8 pixels, SBO = patch row pitch.
          // (64-byte rows: a 32-channel 16-bit pixel IS the operand row, SWIZZLE_64B atoms of 8 rows x 64 B, two MMAs per tap)
          const uint32_t w_addr = smem_u32(smem);
          if (p.patch_layout == UMMA_LAYOUT_NONE) {
            // 16-byte pixels (RGB stored as 8 fp16 channels), UNSWIZZLED operands: a core matrix (8 rows x 16 B) is 8
            // consecutive pixels of a patch row; the K-adjacent core matrix (LBO) is the SAME memory one pixel further, so
            // the horizontal taps of a filter row are the K dimension of the MMA with no replication at all: one MMA
            // (K = 16) = two taps x 8 channels, its 8-row groups (SBO) are the tile's 16 rows.
            mbar_wait(&full_bar[stage], phase, 0x300 + stage, p.err_sink);
            tcgen05_fence_after();
            const uint32_t sbo = p.patch_w * 16;
            const uint64_t da0 = make_smem_desc(w_addr + p.w_total_bytes + stage * p.stage_bytes, 16, sbo, UMMA_LAYOUT_NONE);
            const uint64_t db0 = make_smem_desc(w_addr, p.block_n * 16, 128, UMMA_LAYOUT_NONE);
            const uint32_t b_mma = p.block_n * 32;  // two K chunks of [cout x 16 B]
            uint32_t first = 0, a_off = 0, b_off = 0;
            for (int r = 0; r < p.patch_k; ++r) {
              uint64_t da = desc_adv(da0, a_off), db = desc_adv(db0, b_off);
              for (int j = 0; j < p.patch_kmma; ++j) {
                umma_ss<kTf32>(d_tmem, da, db, idesc, first);
                first = 1;
                da += 2;
                db += b_mma >> 4;
              }
              a_off += sbo;
              b_off += p.w_tile_bytes;
            }
            umma_commit(&empty_bar[stage]);
            if (++stage == p.num_stages) { stage = 0; phase ^= 1; }
            umma_commit(&tmem_full[acc]);
            next_acc();
            continue;
          }
          const uint32_t rb = p.patch_row_bytes;
          const uint32_t sbo = p.patch_w * rb;
          const uint64_t db0 = make_smem_desc(w_addr, 16, 8 * rb, p.patch_layout);
          const bool four = p.patch_kmma == 4;
          uint32_t first = 0;
          for (int c = 0; c < p.patch_nch; ++c) {
            mbar_wait(&full_bar[stage], phase, 0x300 + stage, p.err_sink);
            tcgen05_fence_after();
            const uint64_t da0 = make_smem_desc(w_addr + p.w_total_bytes + stage * p.stage_bytes, 16, sbo, p.patch_layout);
            uint32_t a_off = 0, b_off = c * p.w_tile_bytes;
            const uint32_t b_step = p.patch_nch * p.w_tile_bytes;
            for (int r = 0; r < p.patch_k; ++r) {
              uint32_t a_rs = a_off;
              for (int s_ = 0; s_ < p.patch_k; ++s_) {
                const uint64_t da = desc_adv(da0, a_rs), db = desc_adv(db0, b_off);
                umma_ss<kTf32>(d_tmem, da, db, idesc, first);
                umma_ss<kTf32>(d_tmem, da + 2, db + 2, idesc, 1u);
                if (four) {
                  umma_ss<kTf32>(d_tmem, da + 4, db + 4, idesc, 1u);
                  umma_ss<kTf32>(d_tmem, da + 6, db + 6, idesc, 1u);
                }
                first = 1;
                a_rs += rb;
                b_off += b_step;
              }
              a_off += sbo;
            }
            umma_commit(&empty_bar[stage]);
            if (++stage == p.num_stages) { stage = 0; phase ^= 1; }
          }
        } else {
          const uint64_t d0 = make_smem_desc(smem_u32(smem), 16, 1024, UMMA_LAYOUT_SW128);
          uint32_t first = 0;
          for (int kc = 0; kc < p.num_k_chunks; ++kc) {
            mbar_wait(&full_bar[stage], phase, 0x300 + stage, p.err_sink);
            tcgen05_fence_after();
            const uint64_t da = desc_adv(d0, stage * p.stage_bytes), db = da + (kABytes >> 4);
            umma_ss<kTf32>(d_tmem, da, db, idesc, first);  // 4 x (32 bytes of K) per 128-byte swizzle row
            umma_ss<kTf32>(d_tmem, da + 2, db + 2, idesc, 1u);
            umma_ss<kTf32>(d_tmem, da + 4, db + 4, idesc, 1u);
            umma_ss<kTf32>(d_tmem, da + 6, db + 6, idesc, 1u);
            first = 1;
            umma_commit(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
            if (++stage == p.num_stages) { stage = 0; phase ^= 1; }
          }
        }
        umma_commit(&tmem_full[acc]);  // accumulator complete -> epilogue
        next_acc();
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (8 warps: lane quarter = warp & 3, column half = (warp - 4) >> 2) ==========
    const int q = warp & 3;          // TMEM lane quarter this warp may read
    const int half = (warp - 4) >> 2;  // interleaved 16-column chunks: half, half + 2, ... (kOcc2: one warp per quarter)
    const int m = q * 32 + lane;
    // alpha = (1/sigma) * s_y / (s_x * s_w * s_mul): the accumulator is in stored units of x and w, the output in stored
    // units of y (LeakyReLU / ReLU are positively homogeneous; tanh / sigmoid heads write true values: s_y = s_mul = 1)
    const float bmul = (p.sy ? __ldg(p.sy) : 1.0f) / (p.smul ? __ldg(p.smul) : 1.0f);
    const float alpha = (p.alpha ? __ldg(p.alpha) : 1.0f) * bmul / ((p.sx ? __ldg(p.sx) : 1.0f) * (p.sw ? __ldg(p.sw) : 1.0f));
    // pre-mul output in its own scale: the accumulator times alpha is s_y / s_mul times the true value
    const float pre_scale = p.premul ? (p.spre ? __ldg(p.spre) : 1.0f) / bmul : 1.0f;
    const int mw = m & (p.tw - 1), mh = (m >> p.tw_log2) & (p.th - 1), mn = m >> (p.tw_log2 + p.th_log2);
    int acc = 0;
    uint32_t acc_phase = 0;
    float* stat_slice = s_stats + (warp - 4) * 256;
    int stat_n = -1;
    if (p.in_stats) {
      for (int i = lane; i < 256; i += 32) stat_slice[i] = 0.f;
    }
    // element (slot, lane) of the slice = column (slot*32 + half*16 + (lane & 15)), sum if lane < 16 else sum of squares
    auto flush_stats = [&](int key) {
      const int n_img = key / p.n_tiles, nt = key % p.n_tiles;
      for (int slot = 0; slot * 32 + half * 16 < p.block_n; ++slot) {
        const int col = nt * p.block_n + slot * 32 + half * 16 + (lane & 15);
        const float v = stat_slice[slot * 32 + lane];
        stat_slice[slot * 32 + lane] = 0.f;
        if (col < p.cout && v != 0.f)
          atomicAdd(p.in_stats + ((long long)n_img * p.cout + col) * 2 + (lane >> 4), (double)v);
      }
    };
    TileIter it;
    it.init(p, blockIdx.x, gridDim.x);
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, it.next(p)) {
      const TileCoord tc = it.coord(p);
      if (p.in_stats && tc.n0 * p.n_tiles + tc.nt != stat_n) {
        if (stat_n >= 0) flush_stats(stat_n);
        stat_n = tc.n0 * p.n_tiles + tc.nt;
      }
      const int wo = tc.wo0 + mw, ho = tc.ho0 + mh, n = tc.n0 + mn;
      const bool valid = (wo < p.Wo) && (ho < p.Ho) && (n < p.Nimg);
      mbar_wait(&tmem_full[acc], acc_phase, 0x400 + acc, p.err_sink);
      tcgen05_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * kAccCols;
      const int colbase = tc.nt * p.block_n;
      if (p.out_nchw) {
        if (half == 0) epilogue_planar(p, taddr, colbase, n, ho, wo, valid, alpha);
      } else {
        const long long o = (long long)n * p.out_img + (long long)ho * p.out_row + (long long)wo * p.out_pix + colbase;
        const long long mo = p.mul ? (long long)n * p.mul_img + (long long)ho * p.mul_row + (long long)wo * p.mul_pix + colbase : 0;
        const long long po = p.premul ? (long long)n * p.pre_img + (long long)ho * p.pre_row + (long long)wo * p.pre_pix + colbase : 0;
        switch (p.act) {
          case UEGAN_ACT_LRELU: epilogue_nhwc<UEGAN_ACT_LRELU, kOcc2 ? 1 : 2>(p, taddr, colbase, half, o, mo, valid, alpha, bmul, stat_slice, n, ho, wo, s_bias, po, pre_scale); break;
          case UEGAN_ACT_RELU: epilogue_nhwc<UEGAN_ACT_RELU, kOcc2 ? 1 : 2>(p, taddr, colbase, half, o, mo, valid, alpha, bmul, stat_slice, n, ho, wo, s_bias, po, pre_scale); break;
          case UEGAN_ACT_NONE: epilogue_nhwc<UEGAN_ACT_NONE, kOcc2 ? 1 : 2>(p, taddr, colbase, half, o, mo, valid, alpha, bmul, stat_slice, n, ho, wo, s_bias, po, pre_scale); break;
          default: epilogue_nhwc<-1, kOcc2 ? 1 : 2>(p, taddr, colbase, half, o, mo, valid, alpha, bmul, stat_slice, n, ho, wo, s_bias, po, pre_scale); break;
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (p.in_stats && stat_n >= 0) flush_stats(stat_n);
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 2 * kAccCols);
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
static int pow2_ceil(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

struct PackGeom {
  int chunk_elems, chunks_per_row, row_pad, cout_pad;
};
static PackGeom pack_geom(int cout, int cin_stored, int k, int dtype) {
  PackGeom g;
  g.chunk_elems = 128 / dtype_size(dtype);
  g.chunks_per_row = (k * cin_stored + g.chunk_elems - 1) / g.chunk_elems;
  g.row_pad = g.chunks_per_row * g.chunk_elems;
  g.cout_pad = (cout + 15) / 16 * 16;
  return g;
}

int launch_conv_fprop(const uegan_conv_desc& d, cudaStream_t stream) {
  const uegan_tensor& x = d.x;
  const int es = dtype_size(x.dtype);
  UEGAN_CHECK(dtype_ok(x.dtype), "conv: bad x dtype %d", x.dtype);
  UEGAN_CHECK((x.c * es) % 16 == 0, "conv: x.c*elem (%d) must be a multiple of 16 bytes", x.c * es);
  UEGAN_CHECK(d.stride == 1 || d.stride == 2, "conv: stride %d unsupported", d.stride);
  UEGAN_CHECK(d.pad <= x.halo, "conv: pad %d exceeds input halo %d", d.pad, x.halo);
  UEGAN_CHECK(d.k >= 1 && d.k <= 7, "conv: k %d unsupported", d.k);
  UEGAN_CHECK(d.w_packed != nullptr && x.data != nullptr, "conv: null pointer");
  const int Ho = (x.h + 2 * d.pad - d.k) / d.stride + 1;
  const int Wo = (x.w + 2 * d.pad - d.k) / d.stride + 1;
  UEGAN_CHECK(Ho >= 1 && Wo >= 1, "conv: empty output");
  const PackGeom g = pack_geom(d.cout, x.c, d.k, x.dtype);

  ConvParams p;
  memset(&p, 0, sizeof(p));
  p.tw = pow2_ceil(Wo < 16 ? Wo : 16);
  p.th = pow2_ceil(Ho);
  if (p.th > 128 / p.tw) p.th = 128 / p.tw;
  p.tn = 128 / (p.tw * p.th);
  for (p.tw_log2 = 0; (1 << p.tw_log2) < p.tw; ++p.tw_log2) {}
  for (p.th_log2 = 0; (1 << p.th_log2) < p.th; ++p.th_log2) {}
  p.tiles_w = (Wo + p.tw - 1) / p.tw;
  p.tiles_h = (Ho + p.th - 1) / p.th;
  p.tiles_img = (x.n + p.tn - 1) / p.tn;
  p.block_n = g.cout_pad <= 256 ? g.cout_pad : 256;
  p.n_tiles = (g.cout_pad + p.block_n - 1) / p.block_n;
  {
    // few-tile launches (deep layers at small batch: d5, its data gradient, ...): split N while the doubled tile count
    // still fits one wave -- an N >= 64 MMA costs ~N/2 cycles, so halving N halves a CTA's time and doubles the CTAs
    const char* env = getenv("UEGAN_NO_NSPLIT");
    const int m_tiles = p.tiles_w * p.tiles_h * p.tiles_img;
    if (!(env && env[0] == '1'))
      while (p.block_n >= 128 && p.block_n % 32 == 0 && 2 * m_tiles * p.n_tiles <= num_sms()) {
        p.block_n /= 2;
        p.n_tiles = (g.cout_pad + p.block_n - 1) / p.block_n;
      }
  }
  p.total_tiles = p.tiles_w * p.tiles_h * p.tiles_img * p.n_tiles;
  p.kh = d.k;
  p.chunks_per_row = g.chunks_per_row;
  p.chunk_elems = g.chunk_elems;
  p.num_k_chunks = d.k * g.chunks_per_row;
  p.stage_bytes = kABytes + ((p.block_n * 128 + 1023) / 1024) * 1024;
  // Two CTAs per SM (conv_fprop_kernel<.., .., 1>) for small-N launches with at least two waves of tiles: budget ~98 KB
  bool occ2 = false;
  {
    const char* env = getenv("UEGAN_CONV_OCC");
    const int want = env ? atoi(env) : 2;
    occ2 = want >= 2 && p.block_n <= 128 && !d.in_stats && p.total_tiles >= 2 * num_sms();
  }
  long long budget = occ2 ? 98 * 1024 : 200 * 1024;
  if (occ2 && budget / p.stage_bytes < 3) { occ2 = false; budget = 200 * 1024; }  // (re-checked for the patch modes below)
  p.num_stages = (int)(budget / p.stage_bytes);
  if (p.num_stages > kMaxStages) p.num_stages = kMaxStages;
  p.Wo = Wo; p.Ho = Ho; p.Nimg = x.n; p.cout = d.cout;
  p.act = d.act;
  p.ab_fmt = x.dtype == UEGAN_F32 ? UMMA_TF32 : (x.dtype == UEGAN_BF16 ? UMMA_BF16 : UMMA_F16);
  p.bias = d.bias;
  p.alpha = d.alpha;
  p.sx = x.scale;
  p.sw = d.w_scale;
  p.out_nchw = d.out_nchw;
  p.residual_nchw = d.residual_nchw;
  p.aux_nchw = d.aux_nchw;
  p.err_sink = error_sink_device();
  p.in_stats = d.in_stats;
  const int ymul = d.y_mul > 1 ? d.y_mul : 1;
  if (d.out_nchw) {
    UEGAN_CHECK(d.cout <= 16, "conv: planar output needs cout <= 16 (got %d)", d.cout);
  } else {
    const uegan_tensor& y = d.y;
    UEGAN_CHECK(y.data != nullptr, "conv: null output");
    if (ymul == 1 && d.y_off_h == 0 && d.y_off_w == 0) {
      UEGAN_CHECK(y.n == x.n && y.h == Ho && y.w == Wo, "conv: y is %dx%dx%d, expected %dx%dx%d", y.n, y.h, y.w, x.n,
                  Ho, Wo);
    } else {
      // strided view: output (a, b) lands at (y_off_h + a*y_mul, y_off_w + b*y_mul); rows/cols beyond y are dropped
      UEGAN_CHECK(y.n == x.n && d.y_off_h >= 0 && d.y_off_w >= 0 && d.y_off_h < y.h && d.y_off_w < y.w && !d.mul,
                  "conv: bad strided output view");
      const int hv = (y.h - d.y_off_h + ymul - 1) / ymul, wv = (y.w - d.y_off_w + ymul - 1) / ymul;
      if (p.Ho > hv) p.Ho = hv;
      if (p.Wo > wv) p.Wo = wv;
    }
    UEGAN_CHECK(d.cout % 16 == 0, "conv: NHWC output needs cout %% 16 == 0 (got %d)", d.cout);
    UEGAN_CHECK(!d.bias || d.cout <= 512, "conv: a bias needs cout <= 512 (staged in shared memory; got %d)", d.cout);
    UEGAN_CHECK(d.y_c_off >= 0 && d.y_c_off + (d.y_cls_c > 0 ? d.y_cls_c : d.cout) <= y.c && d.y_c_off % 8 == 0,
                "conv: bad channel slice");
    UEGAN_CHECK((y.c * dtype_size(y.dtype)) % 16 == 0, "conv: y.c misaligned");
    UEGAN_CHECK(dtype_ok(y.dtype), "conv: bad y dtype %d", y.dtype);
    p.out_kind = y.dtype;
    p.sy = y.scale;
    UEGAN_CHECK(!(y.scale || (d.mul && d.mul->scale)) || d.act == UEGAN_ACT_NONE || d.act == UEGAN_ACT_LRELU ||
                    d.act == UEGAN_ACT_RELU,
                "conv: a scaled output needs a positively homogeneous activation (none / LeakyReLU / ReLU)");
    UEGAN_CHECK(!(y.scale && d.in_stats), "conv: in_stats is not available with a scaled output");
    p.out_pix = y.c;
    p.out_row = t_wp(y) * y.c;
    p.out_img = t_hp(y) * p.out_row;
    const long long off = (long long)(y.halo + d.y_off_h) * p.out_row + (long long)(y.halo + d.y_off_w) * p.out_pix +
                          d.y_c_off;
    if (d.y_cls_c > 0) {
      UEGAN_CHECK(ymul == 2 && d.y_off_h == 0 && d.y_off_w == 0 && d.cout == 4 * d.y_cls_c && d.y_cls_c % 16 == 0 &&
                      d.y_c_off + d.y_cls_c <= y.c && !d.mask && !d.mul && !d.bias && !d.in_stats,
                  "conv: merged parity classes need y_mul 2, no offsets, cout = 4 * y_cls_c, no bias / mask / mul / stats");
      p.cls_c = d.y_cls_c; p.cls_h = y.h; p.cls_w = y.w;
      p.cls_row = p.out_row; p.cls_pix = p.out_pix;
    }
    p.out_pix *= ymul;
    p.out_row *= ymul;
    p.out = static_cast<uint8_t*>(y.data) + off * dtype_size(y.dtype);
    if (d.mask) {
      const uegan_tensor& kt = *d.mask;
      UEGAN_CHECK(dtype_ok(kt.dtype) && kt.n == y.n && kt.h == y.h && kt.w == y.w && kt.c >= d.cout && ymul == 1,
                  "conv: mask tensor mismatch");
      p.mask_pix = kt.c;
      p.mask_row = t_wp(kt) * kt.c;
      p.mask_img = t_hp(kt) * p.mask_row;
      const long long koff = (long long)kt.halo * p.mask_row + (long long)kt.halo * p.mask_pix;
      p.mask = static_cast<const uint8_t*>(kt.data) + koff * dtype_size(kt.dtype);
      p.mask_kind = kt.dtype;
      p.mask_act = d.mask_act;
    }
    if (d.mul) {
      const uegan_tensor& mt = *d.mul;
      UEGAN_CHECK(mt.dtype == y.dtype && mt.n == y.n && mt.h == y.h && mt.w == y.w && mt.c >= d.cout,
                  "conv: mul tensor mismatch");
      p.mul_pix = mt.c;
      p.mul_row = t_wp(mt) * mt.c;
      p.mul_img = t_hp(mt) * p.mul_row;
      const long long moff = (long long)mt.halo * p.mul_row + (long long)mt.halo * p.mul_pix;
      p.mul = static_cast<const uint8_t*>(mt.data) + moff * dtype_size(mt.dtype);
      p.smul = mt.scale;
    }
    if (d.y_reflect_halo) {
      UEGAN_CHECK(y.halo >= 1 && y.dtype != UEGAN_F32 && ymul == 1 && d.y_off_h == 0 && d.y_off_w == 0 && !d.y_cls_c &&
                      y.h > 2 * y.halo + 1 && y.w > 2 * y.halo + 1 && !d.out_nchw,
                  "conv: y_reflect_halo needs a dense 16-bit NHWC output with a halo and h, w > 2 * halo + 1");
      p.reflect_halo = y.halo;
    }
    if (d.y_premul) {
      const uegan_tensor& pt = *d.y_premul;
      UEGAN_CHECK(d.mul && pt.data && pt.dtype == y.dtype && y.dtype != UEGAN_F32 && pt.n == y.n && pt.h == y.h && pt.w == y.w &&
                      pt.c >= d.cout && (pt.c * dtype_size(pt.dtype)) % 16 == 0 && !d.in_stats,
                  "conv: y_premul needs `mul`, a 16-bit tensor of y's extent and no in_stats");
      p.pre_pix = pt.c;
      p.pre_row = t_wp(pt) * pt.c;
      p.pre_img = t_hp(pt) * p.pre_row;
      const long long poff = (long long)pt.halo * p.pre_row + (long long)pt.halo * p.pre_pix;
      p.premul = static_cast<uint8_t*>(pt.data) + poff * dtype_size(pt.dtype);
      p.spre = pt.scale;
    }
  }

  // ---- patch mode?  stride 1, k > 1, pixel = whole 128-byte chunks, single N tile, weights fit next to >= 2 A stages
  const CUtensorMapDataType dt = x.dtype == UEGAN_BF16  ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                 : x.dtype == UEGAN_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16
                                                        : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  const uint64_t pix_b = (uint64_t)x.c * es;
  const uint64_t row_b = (uint64_t)t_wp(x) * pix_b;
  const uint64_t img_b = (uint64_t)t_hp(x) * row_b;
  bool patch = false, stream_w = false;
  {
    const char* env = getenv("UEGAN_NO_PATCH");
    const char* env2 = getenv("UEGAN_NO_STREAM");
    const int PH = 16 + d.k - 1;
    int PW = 8 + d.k - 1;
    // 64-byte pixels (32 channels of a 16-bit type): the same patch mode on SWIZZLE_64B rows -- the plain mode's
    // overlapping 128-byte windows re-fetch every input byte ~12 times through L2 and bound those layers (r2p)
    const char* env64 = getenv("UEGAN_NO_PATCH64");
    const bool row64 = pix_b == 64 && es == 2 && !(env64 && env64[0] == '1');
    // 16-byte pixels (RGB as 8 channels of a 16-bit type): unswizzled operands, the taps of a filter row are the MMA's K
    const char* env16 = getenv("UEGAN_NO_PATCH16");
    const bool row16 = pix_b == 16 && es == 2 && !(env16 && env16[0] == '1');
    const int kpad = (d.k + 1) / 2 * 2;  // taps per filter row, rounded to whole K = 16 MMAs (the extra tap has zero weights)
    const int rb = row16 ? 16 : (row64 ? 64 : 128);
    if (row16) PW = 8 + kpad - 1;
    const long long a_stage = ((long long)PH * PW * rb + 1023) / 1024 * 1024;
    const long long w_tile = row16 ? (long long)kpad * p.block_n * 16 : (long long)p.block_n * rb;
    const long long w_total = row16 ? (w_tile * d.k + 1023) / 1024 * 1024 : w_tile * d.k * d.k * (row64 ? 1 : pix_b / 128);
    const bool shape_ok = !(env && env[0] == '1') && d.stride == 1 && d.k > 1 && (pix_b % 128 == 0 || row64 || row16) &&
                          Ho >= 16 && Wo >= 8 && (!row64 || w_total % 1024 == 0) && (!row16 || g.cout_pad <= 256);
    bool resident = shape_ok && p.n_tiles == 1 && w_total + 2 * a_stage <= budget;
    if (occ2 && shape_ok && p.n_tiles == 1 && w_total + 2 * a_stage > budget && w_total + 2 * a_stage <= 200 * 1024) {
      // the resident-weight patch mode needs the whole SM's shared memory here: it beats plain mode at two CTAs per SM
      // (two patch stages per CTA are enough at two CTAs per SM: dec4 forward 0.274 -> 0.219 ms, r3c)
      occ2 = false;
      budget = 200 * 1024;
      p.num_stages = (int)(budget / p.stage_bytes);
      if (p.num_stages > kMaxStages) p.num_stages = kMaxStages;
      resident = true;
    }
    const bool resident_deep = resident && w_total + 3 * a_stage <= budget;  // >= 3 patch stages in flight
    // Weights that do not fit (or leave only two patch stages: the TMA latency of a patch is then exposed) are streamed
    // through their own ring: the A operand is still fetched once per tile instead of once per tap.
    // (N = 256 launches are MMA-bound in plain mode already: 85-97 % of the tensor peak; they stay there.)
    const char* env3 = getenv("UEGAN_STREAM_MAXN");
    const int stream_max_n = env3 ? atoi(env3) : 0;  // opt-in: measured 1-6 % SLOWER than plain mode
    if (shape_ok && !occ2 && !row64 && !row16 && !resident_deep && !(env2 && env2[0] == '1') && p.block_n <= stream_max_n && w_tile % 1024 == 0 &&
        3 * a_stage + 4 * w_tile <= budget) {
      stream_w = true;
      p.tw = 8; p.th = 16; p.tn = 1; p.tw_log2 = 3; p.th_log2 = 4;
      p.tiles_w = (Wo + 7) / 8;
      p.tiles_h = (Ho + 15) / 16;
      p.tiles_img = x.n;
      p.total_tiles = p.tiles_w * p.tiles_h * p.tiles_img * p.n_tiles;
      p.patch_w = PW;
      p.patch_nch = (int)(pix_b / 128);
      p.patch_k = d.k;
      p.patch_off = x.halo - d.pad;
      p.w_tile_bytes = (int)w_tile;
      p.w_total_bytes = 0;
      p.stage_bytes = (int)a_stage;
      p.stage_tx_bytes = PH * PW * 128;
      p.num_stages = 3;
      p.b_ring_off = p.num_stages * p.stage_bytes;
      p.b_slots = (int)((budget - p.b_ring_off) / w_tile);
      if (p.b_slots > kMaxBSlots) p.b_slots = kMaxBSlots;
    } else if (resident) {
      patch = true;
      p.tw = 8; p.th = 16; p.tn = 1; p.tw_log2 = 3; p.th_log2 = 4;
      p.tiles_w = (Wo + 7) / 8;
      p.tiles_h = (Ho + 15) / 16;
      p.tiles_img = x.n;
      p.total_tiles = p.tiles_w * p.tiles_h * p.tiles_img;
      p.patch_w = PW;
      p.patch_nch = (row64 || row16) ? 1 : (int)(pix_b / 128);
      p.patch_k = d.k;
      p.patch_off = x.halo - d.pad;
      p.patch_row_bytes = rb;
      p.patch_kmma = row16 ? kpad / 2 : rb / 32;
      p.patch_layout = row16 ? UMMA_LAYOUT_NONE : (row64 ? UMMA_LAYOUT_SW64 : UMMA_LAYOUT_SW128);
      p.patch_ce = rb / es;
      p.patch_cpr = g.row_pad / p.patch_ce;
      p.w_tile_bytes = (int)w_tile;
      p.w_total_bytes = (int)w_total;
      p.stage_bytes = (int)a_stage;
      p.stage_tx_bytes = PH * PW * rb;
      p.num_stages = (int)((budget - w_total) / a_stage);
      if (p.num_stages > kMaxStages) p.num_stages = kMaxStages;
    }
  }
  // ---- tensor maps
  CUtensorMap tmA, tmB;
  const bool patch64 = patch && p.patch_row_bytes == 64, patch16 = patch && p.patch_row_bytes == 16;
  if (patch || stream_w) {
    uint64_t dims[4] = {(uint64_t)x.c, (uint64_t)t_wp(x), (uint64_t)t_hp(x), (uint64_t)x.n};
    uint64_t strides[3] = {pix_b, row_b, img_b};
    uint32_t box[4] = {(uint32_t)(patch16 ? 8 : (patch64 ? 32 : g.chunk_elems)), (uint32_t)p.patch_w,
                       (uint32_t)(16 + d.k - 1), 1u};
    if (encode_tiled(&tmA, dt, 4, x.data, dims, strides, box,
                     patch16 ? CU_TENSOR_MAP_SWIZZLE_NONE
                             : (patch64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B)))
      return -1;
  } else {
    uint8_t* a_base = static_cast<uint8_t*>(x.data) + (uint64_t)(x.halo - d.pad) * row_b + (uint64_t)(x.halo - d.pad) * pix_b;
    uint64_t dims[5] = {(uint64_t)g.row_pad, (uint64_t)Wo, (uint64_t)d.k, (uint64_t)Ho, (uint64_t)x.n};
    uint64_t strides[4] = {(uint64_t)d.stride * pix_b, row_b, (uint64_t)d.stride * row_b, img_b};
    uint32_t box[5] = {(uint32_t)g.chunk_elems, (uint32_t)p.tw, 1u, (uint32_t)p.th, (uint32_t)p.tn};
    if (encode_tiled(&tmA, dt, 5, a_base, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return -1;
  }
  if (patch16) {
    // weights [cout_pad][k][row_pad] read as {8 values, cout, K chunk, filter row}: the box {8, cout, kpad / .., 1} lands as
    // [K chunk][cout][8 values] -- core matrices of 8 output channels x 16 B, K-adjacent ones cout * 16 B apart
    const uint64_t ktot = (uint64_t)d.k * g.row_pad;
    uint64_t dims[4] = {8u, (uint64_t)g.cout_pad, (uint64_t)(g.row_pad / 8), (uint64_t)d.k};
    uint64_t strides[3] = {ktot * es, 16u, (uint64_t)g.row_pad * es};
    uint32_t box[4] = {8u, (uint32_t)p.block_n, (uint32_t)(2 * p.patch_kmma), 1u};
    if (encode_tiled(&tmB, dt, 4, const_cast<void*>(d.w_packed), dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE)) return -1;
  } else {
    const uint64_t ktot = (uint64_t)d.k * g.row_pad;
    uint64_t dims[2] = {ktot, (uint64_t)g.cout_pad};
    uint64_t strides[1] = {ktot * es};
    uint32_t box[2] = {(uint32_t)(patch64 ? 32 : g.chunk_elems), (uint32_t)p.block_n};
    if (encode_tiled(&tmB, dt, 2, const_cast<void*>(d.w_packed), dims, strides, box,
                     patch64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B))
      return -1;
  }
  if (d.in_stats) {
    UEGAN_CHECK(!d.out_nchw && !d.mul, "conv: in_stats needs a plain NHWC epilogue");
    UEGAN_CHECK(p.tn == 1, "conv: in_stats needs tiles within one image (Ho*Wo >= 128); use uegan_instance_norm");
    UEGAN_CUDA(cudaMemsetAsync(d.in_stats, 0, sizeof(double) * 2 * (size_t)x.n * d.cout, stream));
  }
  {
    // two MMA issuers when a tile's MMAs are cheap (N <= 64: the issuing thread's instruction stream is the bottleneck)
    const char* env = getenv("UEGAN_ISSUERS");
    const int per_tile = patch ? p.patch_nch : p.num_k_chunks;
    p.n_issuers = env ? atoi(env) : (p.block_n <= 64 ? 2 : 1);
    // An issuer SKIPS the other's stages without waiting on them, and mbarrier parity waits only tell phases apart that
    // are at most one ring pass away: the ring must hold two whole tiles, or a skipped-ahead wait would pass on a stale
    // phase (measured: launch failure with 72-chunk tiles on an 8-stage ring).
    if (stream_w || p.n_issuers != 2 || p.num_stages < 2 * per_tile) p.n_issuers = 1;
  }
  const int smem_bytes = stream_w ? p.b_ring_off + p.b_slots * p.w_tile_bytes + 1024
                                  : (patch ? p.w_total_bytes : 0) + p.num_stages * p.stage_bytes + 1024;
  const int max_ctas = occ2 ? 2 * num_sms() : num_sms();
  int grid = p.total_tiles < max_ctas ? p.total_tiles : max_ctas;
  {
    static bool attr_set = false;
    if (!attr_set) {
      UEGAN_CUDA(cudaFuncSetAttribute(conv_fprop_kernel<1, 0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024));
      UEGAN_CUDA(cudaFuncSetAttribute(conv_fprop_kernel<0, 0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024));
      UEGAN_CUDA(cudaFuncSetAttribute(conv_fprop_kernel<1, 1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024));
      UEGAN_CUDA(cudaFuncSetAttribute(conv_fprop_kernel<0, 1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024));
      UEGAN_CUDA(cudaFuncSetAttribute(conv_fprop_kernel<1, 2, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024));
      UEGAN_CUDA(cudaFuncSetAttribute(conv_fprop_kernel<0, 2, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024));
      UEGAN_CUDA(cudaFuncSetAttribute(conv_fprop_kernel<1, 0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
      UEGAN_CUDA(cudaFuncSetAttribute(conv_fprop_kernel<0, 0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
      UEGAN_CUDA(cudaFuncSetAttribute(conv_fprop_kernel<1, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
      UEGAN_CUDA(cudaFuncSetAttribute(conv_fprop_kernel<0, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
      attr_set = true;
    }
  }
  UEGAN_CHECK(!occ2 || smem_bytes <= 100 * 1024, "conv: internal: two-CTA build with %d bytes of shared memory", smem_bytes);
  if (occ2) {
    if (x.dtype == UEGAN_F32) {
      if (patch) launch_pdl(conv_fprop_kernel<1, 1, 1>, grid, 256, smem_bytes, stream, tmA, tmB, p);
      else launch_pdl(conv_fprop_kernel<1, 0, 1>, grid, 256, smem_bytes, stream, tmA, tmB, p);
    } else {
      if (patch) launch_pdl(conv_fprop_kernel<0, 1, 1>, grid, 256, smem_bytes, stream, tmA, tmB, p);
      else launch_pdl(conv_fprop_kernel<0, 0, 1>, grid, 256, smem_bytes, stream, tmA, tmB, p);
    }
  } else if (x.dtype == UEGAN_F32) {
    if (stream_w) launch_pdl(conv_fprop_kernel<1, 2, 0>, grid, 384, smem_bytes, stream, tmA, tmB, p);
    else if (patch) launch_pdl(conv_fprop_kernel<1, 1, 0>, grid, 384, smem_bytes, stream, tmA, tmB, p);
    else launch_pdl(conv_fprop_kernel<1, 0, 0>, grid, 384, smem_bytes, stream, tmA, tmB, p);
  } else {
    if (stream_w) launch_pdl(conv_fprop_kernel<0, 2, 0>, grid, 384, smem_bytes, stream, tmA, tmB, p);
    else if (patch) launch_pdl(conv_fprop_kernel<0, 1, 0>, grid, 384, smem_bytes, stream, tmA, tmB, p);
    else launch_pdl(conv_fprop_kernel<0, 0, 0>, grid, 384, smem_bytes, stream, tmA, tmB, p);
  }
  UEGAN_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------
// weight packing: OIHW fp32 -> [cout_pad][k][row_pad] (row = (s, c) with c over the STORED channels of x)
// ------------------------------------------------------------------------------------------
// mode 0: fprop operand.  out[o][r][(s, c)] = w[o][cin_first + c][r][s]
// mode 1: dgrad operand of a stride-`q` conv, parity class (pi, pj), packed kernel size k (= ceil(k_orig / q)):
//         out[o][t_r][(t_s, c)] = w[c][cin_first + o][q*(k-1-t_r) + pi][q*(k-1-t_s) + pj]   (0 beyond k_orig)
//         i.e. the roles of O and I are swapped and the taps run backwards (transposed convolution).
template <typename T>
__global__ void pack_weight_kernel(const float* __restrict__ w, T* __restrict__ out, int cout, int cin_total,
                                   int cin_first, int cin, int cin_stored, int k, int row_pad, int cout_pad,
                                   int mode, int k_orig, int q, int pi, int pj, long long total,
                                   const float* __restrict__ wscale) {
  pdl_sync();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  if (gridDim.y > 1) {  // all four parity classes of a stride-2 data-gradient operand in one launch: class = blockIdx.y
    pi = blockIdx.y >> 1;
    pj = blockIdx.y & 1;
    out += (long long)blockIdx.y * total;
  }
  const int x = (int)(i % row_pad);
  const int r = (int)((i / row_pad) % k);
  const int o = (int)(i / ((long long)row_pad * k));
  float v = 0.f;
  const int s = x / cin_stored, c = x % cin_stored;
  if (o < cout && s < k && c < cin) {
    if (mode == 0) {
      v = w[(((long long)o * cin_total + cin_first + c) * k + r) * k + s];
    } else {
      const int rr = q * (k - 1 - r) + pi, ss = q * (k - 1 - s) + pj;
      if (rr < k_orig && ss < k_orig)
        v = w[(((long long)c * cin_total + cin_first + o) * k_orig + rr) * k_orig + ss];
    }
  }
  if (wscale) v *= __ldg(wscale);
  if constexpr (sizeof(T) == 4) {
    out[i] = round_tf32(v);
  } else if constexpr (std::is_same<T, __half>::value) {
    out[i] = __float2half_rn(v);
  } else {
    out[i] = __float2bfloat16_rn(v);
  }
}

}  // namespace uegan

using namespace uegan;

extern "C" {

size_t uegan_packed_weight_bytes(int32_t cout, int32_t cin_stored, int32_t k, int32_t dtype) {
  const PackGeom g = pack_geom(cout, cin_stored, k, dtype);
  return (size_t)g.cout_pad * k * g.row_pad * dtype_size(dtype);
}

static int pack_impl(const float* w_oihw, void* w_packed, int cout, int cin_total, int cin_first, int cin,
                     int cin_stored, int k, int dtype, int mode, int k_orig, int q, int pi, int pj, void* stream,
                     const float* wscale = nullptr, int classes = 1) {
  UEGAN_CHECK(w_oihw && w_packed, "pack_conv_weight: null pointer");
  UEGAN_CHECK(cin <= cin_stored, "pack_conv_weight: cin %d > stored %d", cin, cin_stored);
  UEGAN_CHECK(dtype_ok(dtype), "pack_conv_weight: bad dtype");
  const PackGeom g = pack_geom(cout, cin_stored, k, dtype);
  const long long total = (long long)g.cout_pad * k * g.row_pad;
  const int threads = 256;
  const dim3 blocks((unsigned)((total + threads - 1) / threads), (unsigned)classes);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == UEGAN_F32)
    launch_pdl(pack_weight_kernel<float>, blocks, threads, 0, st, w_oihw, static_cast<float*>(w_packed), cout, cin_total,
                                                          cin_first, cin, cin_stored, k, g.row_pad, g.cout_pad, mode,
                                                          k_orig, q, pi, pj, total, wscale);
  else if (dtype == UEGAN_BF16)
    launch_pdl(pack_weight_kernel<__nv_bfloat16>, blocks, threads, 0, st, w_oihw, static_cast<__nv_bfloat16*>(w_packed), cout,
                                                                   cin_total, cin_first, cin, cin_stored, k, g.row_pad,
                                                                   g.cout_pad, mode, k_orig, q, pi, pj, total, wscale);
  else
    launch_pdl(pack_weight_kernel<__half>, blocks, threads, 0, st, w_oihw, static_cast<__half*>(w_packed), cout, cin_total,
                                                           cin_first, cin, cin_stored, k, g.row_pad, g.cout_pad, mode,
                                                           k_orig, q, pi, pj, total, wscale);
  UEGAN_CUDA(cudaGetLastError());
  return 0;
}

int uegan_pack_conv_weight(const float* w_oihw, void* w_packed, int32_t cout, int32_t cin_total, int32_t cin_first,
                           int32_t cin, int32_t cin_stored, int32_t k, int32_t dtype, int32_t transpose_flip,
                           void* stream) {
  if (!transpose_flip) UEGAN_CHECK(cin_first + cin <= cin_total, "pack_conv_weight: channel range out of bounds");
  return pack_impl(w_oihw, w_packed, cout, cin_total, cin_first, cin, cin_stored, k, dtype, transpose_flip ? 1 : 0, k,
                   1, 0, 0, stream);
}

int uegan_pack_conv_weight_dgrad(const float* w_oihw, void* w_packed, int32_t cout_orig, int32_t cin_total,
                                 int32_t cin_first, int32_t cin, int32_t cout_stored, int32_t k_orig, int32_t stride,
                                 int32_t pi, int32_t pj, int32_t dtype, void* stream) {
  UEGAN_CHECK(stride == 1 || stride == 2, "pack_conv_weight_dgrad: stride %d", stride);
  UEGAN_CHECK(cin_first + cin <= cin_total, "pack_conv_weight_dgrad: channel range out of bounds");
  const int k = (k_orig + stride - 1) / stride;
  // dgrad GEMM: "output" channels = the original input channels [cin_first, cin_first+cin), reduction over the
  // original output channels (stored count cout_stored in the dz tensor)
  return pack_impl(w_oihw, w_packed, cin, cin_total, cin_first, cout_orig, cout_stored, k, dtype, 1, k_orig, stride, pi,
                   pj, stream);
}

int uegan_pack_conv_weight_scaled(const float* w_oihw, void* w_packed, int32_t cout, int32_t cin_total, int32_t cin_first,
                                  int32_t cin, int32_t cin_stored, int32_t k, int32_t dtype, const float* w_scale_dev,
                                  void* stream) {
  UEGAN_CHECK(cin_first + cin <= cin_total, "pack_conv_weight_scaled: channel range out of bounds");
  return pack_impl(w_oihw, w_packed, cout, cin_total, cin_first, cin, cin_stored, k, dtype, 0, k, 1, 0, 0, stream,
                   w_scale_dev);
}

int uegan_pack_conv_weight_dgrad_scaled(const float* w_oihw, void* w_packed, int32_t cout_orig, int32_t cin_total,
                                        int32_t cin_first, int32_t cin, int32_t cout_stored, int32_t k_orig,
                                        int32_t stride, int32_t pi, int32_t pj, int32_t dtype, const float* w_scale_dev,
                                        void* stream) {
  UEGAN_CHECK(stride == 1 || stride == 2, "pack_conv_weight_dgrad_scaled: stride %d", stride);
  UEGAN_CHECK(cin_first + cin <= cin_total, "pack_conv_weight_dgrad_scaled: channel range out of bounds");
  const int k = (k_orig + stride - 1) / stride;
  return pack_impl(w_oihw, w_packed, cin, cin_total, cin_first, cout_orig, cout_stored, k, dtype, 1, k_orig, stride, pi,
                   pj, stream, w_scale_dev);
}

int uegan_pack_conv_weight_dgrad4(const float* w_oihw, void* w_packed, int32_t cout_orig, int32_t cin_total, int32_t cin_first,
                                  int32_t cin, int32_t cout_stored, int32_t k_orig, int32_t dtype, const float* w_scale_dev,
                                  void* stream) {
  UEGAN_CHECK(cin_first + cin <= cin_total, "pack_conv_weight_dgrad4: channel range out of bounds");
  const int k = (k_orig + 1) / 2;
  return pack_impl(w_oihw, w_packed, cin, cin_total, cin_first, cout_orig, cout_stored, k, dtype, 1, k_orig, 2, 0, 0, stream,
                   w_scale_dev, 4);
}

int uegan_conv2d_fprop(const uegan_conv_desc* desc, void* stream) {
  UEGAN_CHECK(desc != nullptr, "conv2d_fprop: null desc");
  return launch_conv_fprop(*desc, static_cast<cudaStream_t>(stream));
}

int uegan_device_error(void) {
  cudaError_t e = cudaDeviceSynchronize();
  unsigned int* sink = error_sink_host();
  const unsigned int v = sink ? *sink : 0;
  if (sink) *sink = 0;
  if (e != cudaSuccess) {
    set_error("device error: %s (watchdog code 0x%x)", cudaGetErrorString(e), v);
    return v ? (int)v : -1;
  }
  return (int)v;
}

}  // extern "C"
