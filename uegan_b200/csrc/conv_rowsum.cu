// "Row-sum" implicit-GEMM convolution for TINY output-channel counts (G's last conv 32 -> 3, k7, models.py:34-35;
// the Discriminator's prediction heads C -> 1, k7 / k5, models.py:110-126,170-182), stride 1, tf32.
//
// The plain / patch kernels of conv_fprop.cu issue one MMA group per filter tap with N = 16 (3 or 1 real columns):
// k*k*C/8 MMAs per 128 output pixels, each bound by its 128-row A fetch -- 1.5 MMAs per output pixel for k7, C = 32.
// Here the HORIZONTAL taps become GEMM columns and only the vertical taps stay in the K loop:
//     D[p][(s, o)] = sum_r sum_c  X[row(p) + r][col(p)][c] * w[o][c][r][s]          N = k*Cout (21 -> 32, 7 -> 16)
//     y[i][x][o]   = sum_s D[(i, x + s)][(s, o)]                                      (epilogue, warp shuffles)
// One staged patch = (TH + k - 1) rows x 32 pixels x 32 channels (one 128-byte row per pixel, SWIZZLE_128B, a single
// 4-D TMA box); an M tile = 4 consecutive patch rows = 128 contiguous smem rows, i.e. the canonical K-major operand,
// and vertical tap r is the same patch read through a descriptor advanced by r*32 rows (4096 B: swizzle phase
// unchanged).  Each TMEM lane quarter (one epilogue warp) holds one patch row of 32 pixels, so the horizontal shift
// x + s is a __shfl_down; a tile yields (32 - k + 1) x TH outputs.  MMAs per output pixel: k*C/8 * 128/(32-k+1)/32
// = 0.27 for k7, C = 32 (5.7x fewer), and the patch is fetched once (1.7x halo overhead) instead of once per tap.
// Warp roles and barriers as in conv_fprop_kernel: warp 0 TMA, warp 1 MMA issue, warp 2 TMEM, warps 4-11 epilogue.
#include "common.cuh"
#include "host_util.h"

namespace uegan {

constexpr int kRsStages = 4;

struct RowsumParams {
  int k, cout, nb;        // nb: N of the MMA = k*cout rounded up to 16
  int nch, cs, cb;        // 128-byte channel chunks per pixel, stored channels, channels per chunk (32 tf32 / 64 fp16)
  int csp;                // nch * cb: channels per pixel of the packed weight rows (zero beyond cs)
  const float *sx, *sw;   // per-tensor scales of x and of the packed weight (NULL = 1)
  int th, msub, two;      // output rows per tile (16 / 8), M tiles per patch (th / 4), output columns per tile
  int tiles_w, tiles_h, total_tiles;
  int patch_off, ph;
  int stage_bytes, num_stages, stage_tx;
  int w_tile_bytes, w_total_bytes;
  int row_bytes;          // one patch row: 32 pixels x 128 B, or x 64 B (32-channel fp16 pixels: SWIZZLE_64B rows, row64)
  int row64;
  int w_tx_bytes;         // exact bytes of the resident weights (w_total_bytes is their 1 KB-rounded smem reservation)
  int occ2;               // planned for two CTAs per SM
  int Wo, Ho, act;
  const float* bias;
  const float* alpha;
  float* out_nchw;
  const float* residual_nchw;
  float* aux_nchw;
  unsigned int* err_sink;
};

__device__ __forceinline__ float rs_act(float v, int act) {
  switch (act) {
    case UEGAN_ACT_LRELU: return v > 0.f ? v : 0.2f * v;
    case UEGAN_ACT_RELU: return fmaxf(v, 0.f);
    case UEGAN_ACT_TANH: return tanhf(v);
    case UEGAN_ACT_SIGMOID: return 1.f / (1.f + __expf(-v));
    default: return v;
  }
}

// one warp, one M tile: lane = pixel column of patch row (4j + q); y[x][o] = sum_s D[x + s][s*COUT + o]
template <int KS, int COUT>
__device__ __forceinline__ void rowsum_epilogue(const RowsumParams& p, uint32_t taddr, int n, int ho, int wo, bool valid,
                                                float alpha) {
  constexpr int NV = KS * COUT;
  uint32_t rr[32];
  {
    uint32_t lo[16];
    tmem_ld16(taddr, lo);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 16; ++i) rr[i] = lo[i];
  }
  if constexpr (NV > 16) {
    uint32_t hi[16];
    tmem_ld16(taddr + 16, hi);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 16; ++i) rr[16 + i] = hi[i];
  }
  float acc[COUT];
#pragma unroll
  for (int o = 0; o < COUT; ++o) acc[o] = __uint_as_float(rr[o]);
#pragma unroll
  for (int s = 1; s < KS; ++s)
#pragma unroll
    for (int o = 0; o < COUT; ++o) acc[o] += __shfl_down_sync(0xffffffffu, __uint_as_float(rr[s * COUT + o]), s);
  if (!valid) return;
  const long long plane = (long long)p.Ho * p.Wo;
#pragma unroll
  for (int o = 0; o < COUT; ++o) {
    const long long off = ((long long)n * COUT + o) * plane + (long long)ho * p.Wo + wo;
    float x = acc[o] * alpha;
    if (p.bias) x += __ldg(p.bias + o);
    x = rs_act(x, p.act);
    if (p.aux_nchw) p.aux_nchw[off] = x;
    if (p.residual_nchw) x = fminf(fmaxf(x + __ldg(p.residual_nchw + off), -1.f), 1.f);
    p.out_nchw[off] = x;
  }
}

template <int kF16>
__global__ void __launch_bounds__(384, 1)
conv_rowsum_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                   const RowsumParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t full_bar[kRsStages];
  __shared__ __align__(8) uint64_t empty_bar[kRsStages];
  __shared__ __align__(8) uint64_t tmem_full[2];
  __shared__ __align__(8) uint64_t tmem_empty[2];
  __shared__ __align__(8) uint64_t w_full;
  __shared__ uint32_t tmem_base_smem;

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
      mbar_init(&tmem_empty[i], 8);
    }
    mbar_init(&w_full, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(&tmem_base_smem, 256);  // 2 accumulator stages x (4 M tiles x 32 columns)
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  pdl_sync();  // everything above is on-chip set-up: it overlaps the tail of the preceding kernel
  const int tiles_per_img = p.tiles_w * p.tiles_h;

  if (warp == 0) {
    if (elect_one()) {
      // ===================== TMA producer =====================
      mbar_arrive_expect_tx(&w_full, (uint32_t)p.w_tx_bytes);
      for (int wi = 0; wi < p.k * p.nch; ++wi) {  // weight tile (r, chunk): [nb rows x 128 B], resident
        const int r = wi / p.nch, c = wi % p.nch;
        tma_load_2d(&tmB, &w_full, smem + wi * p.w_tile_bytes, r * p.csp + c * p.cb, 0);
      }
      uint8_t* sa0 = smem + p.w_total_bytes;
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
        const int n = t / tiles_per_img, tt = t % tiles_per_img;
        const int wo0 = (tt % p.tiles_w) * p.two, ho0 = (tt / p.tiles_w) * p.th;
        for (int c = 0; c < p.nch; ++c) {
          mbar_wait(&empty_bar[stage], phase ^ 1, 0xD00 + stage, p.err_sink);
          mbar_arrive_expect_tx(&full_bar[stage], (uint32_t)p.stage_tx);
          tma_load_4d(&tmA, &full_bar[stage], sa0 + stage * p.stage_bytes, c * p.cb, wo0 + p.patch_off, ho0 + p.patch_off, n);
          if (++stage == p.num_stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      // ===================== MMA issuer =====================
      const uint32_t idesc = make_instr_desc(kF16 ? UMMA_F16 : UMMA_TF32, 128, p.nb);
      const uint32_t w_addr = smem_u32(smem);
      const uint32_t lay = p.row64 ? UMMA_LAYOUT_SW64 : UMMA_LAYOUT_SW128;
      const uint32_t sbo = p.row64 ? 512 : 1024;
      const uint64_t db0 = make_smem_desc(w_addr, 16, sbo, lay);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      mbar_wait(&w_full, 0, 0xE00, p.err_sink);
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1, 0xE10 + acc, p.err_sink);
        tcgen05_fence_after();
        for (int c = 0; c < p.nch; ++c) {
          mbar_wait(&full_bar[stage], phase, 0xE20 + stage, p.err_sink);
          tcgen05_fence_after();
          const uint64_t da0 = make_smem_desc(w_addr + p.w_total_bytes + stage * p.stage_bytes, 16, sbo, lay);
          for (int j = 0; j < p.msub; ++j) {
            const uint32_t d_tmem = tmem_base + acc * 128 + j * 32;
            uint64_t da = desc_adv(da0, (uint32_t)(4 * j) * p.row_bytes);
            uint64_t db = desc_adv(db0, (uint32_t)c * p.w_tile_bytes);
            const uint32_t b_step = (uint32_t)(p.nch * p.w_tile_bytes) >> 4;
            for (int r = 0; r < p.k; ++r) {
              // 4 MMAs per 128-byte row (2 per 64-byte row): 32 bytes of K each (8 tf32 / 16 fp16 values)
              umma_ss<kF16 ? 0 : 1>(d_tmem, da, db, idesc, (c | r) != 0 ? 1u : 0u);
              umma_ss<kF16 ? 0 : 1>(d_tmem, da + 2, db + 2, idesc, 1u);
              if (!p.row64) {
                umma_ss<kF16 ? 0 : 1>(d_tmem, da + 4, db + 4, idesc, 1u);
                umma_ss<kF16 ? 0 : 1>(d_tmem, da + 6, db + 6, idesc, 1u);
              }
              da += p.row_bytes >> 4;
              db += b_step;
            }
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == p.num_stages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tmem_full[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue: lane quarter q = patch row within the M tile, half = M tiles {half, half + 2} ====
    const int q = warp & 3;
    const int half = (warp - 4) >> 2;
    // planar fp32 outputs are true values: the accumulator is in stored units of x and w
    const float alpha = (p.alpha ? __ldg(p.alpha) : 1.0f) / ((p.sx ? __ldg(p.sx) : 1.0f) * (p.sw ? __ldg(p.sw) : 1.0f));
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
      const int n = t / tiles_per_img, tt = t % tiles_per_img;
      const int wo = (tt % p.tiles_w) * p.two + lane, ho0 = (tt / p.tiles_w) * p.th;
      mbar_wait(&tmem_full[acc], acc_phase, 0xF00 + acc, p.err_sink);
      tcgen05_fence_after();
      for (int j = half; j < p.msub; j += 2) {
        const int ho = ho0 + 4 * j + q;
        const bool valid = lane < p.two && wo < p.Wo && ho < p.Ho;
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * 128 + j * 32;
        if (p.k == 7 && p.cout == 3) rowsum_epilogue<7, 3>(p, taddr, n, ho, wo, valid, alpha);
        else if (p.k == 7) rowsum_epilogue<7, 1>(p, taddr, n, ho, wo, valid, alpha);
        else if (p.k == 5) rowsum_epilogue<5, 1>(p, taddr, n, ho, wo, valid, alpha);
        else rowsum_epilogue<3, 1>(p, taddr, n, ho, wo, valid, alpha);
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 256);
}

// out[n][r][c] (n = s*cout + o, nb rows; k*csp columns, csp = stored channels padded to whole 128-byte chunks)
//   = round(w[o][cin_first + c][r][s] * (*wscale)), zero elsewhere; tf32-rounded fp32 or fp16
template <typename T>
__global__ void pack_weight_rowsum_kernel(const float* __restrict__ w, T* __restrict__ out, int cout, int cin_total,
                                          int cin_first, int cin, int csp, int k, int nb, const float* __restrict__ wscale) {
  pdl_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nb * k * csp) return;
  const int c = i % csp, r = (i / csp) % k, nrow = i / (csp * k);
  const int s = nrow / cout, o = nrow % cout;
  float v = 0.f;
  if (s < k && c < cin) v = w[(((long long)o * cin_total + cin_first + c) * k + r) * k + s];
  if (wscale) v *= __ldg(wscale);
  if constexpr (sizeof(T) == 4) {
    uint32_t t;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v));
    out[i] = __uint_as_float(t);
  } else {
    out[i] = __float2half_rn(v);
  }
}

static bool rowsum_shape_ok(int cout, int k) {
  return (k == 7 && (cout == 3 || cout == 1)) || ((k == 5 || k == 3) && cout == 1);
}
static int rowsum_nb(int cout, int k) { return (k * cout + 15) / 16 * 16; }

static int rowsum_cb(int dtype) { return dtype == UEGAN_F32 ? 32 : 64; }
// picks th (16, then 8) so that the resident weights and >= 2 patch stages fit; 0 if neither does
static int rowsum_plan(int cout, int k, int cs, int dtype, RowsumParams* p) {
  const int cb = rowsum_cb(dtype);
  const int nb = rowsum_nb(cout, k), nch = (cs + cb - 1) / cb;
  // 32-channel fp16 pixels are 64 bytes: SWIZZLE_64B rows (half the smem, TMA bytes and MMAs of a half-empty 128-byte row)
  const char* e64 = getenv("UEGAN_NO_ROWSUM64");
  const bool row64 = dtype == UEGAN_F16 && cs == 32 && !(e64 && e64[0] == '1');
  const int rb = row64 ? 64 : 128;
  const long long w_total = ((long long)k * nch * nb * rb + 1023) / 1024 * 1024;
  const char* eocc = getenv("UEGAN_ROWSUM_OCC");
  const int occ = eocc ? atoi(eocc) : 2;
  // first choice: two CTAs per SM (<= 98 KB each: th = 8, >= 3 stages); else one CTA with the whole shared memory
  for (int pass = (occ >= 2 ? 0 : 1); pass < 2; ++pass) {
    const long long budget = pass == 0 ? 98 * 1024 : 200 * 1024;
    for (int th = (pass == 0 ? 8 : 16); th >= 8; th -= 8) {
      const long long stage = (long long)(th + k - 1) * 32 * rb;
      const long long room = budget - w_total;
      if (room >= (pass == 0 ? 3 : 2) * stage) {
        if (p) {
          p->th = th; p->msub = th / 4; p->ph = th + k - 1;
          p->stage_bytes = (int)stage; p->stage_tx = (int)stage;
          p->num_stages = (int)(room / stage) < kRsStages ? (int)(room / stage) : kRsStages;
          p->w_tile_bytes = nb * rb; p->w_total_bytes = (int)w_total;
          p->row_bytes = 32 * rb; p->row64 = row64 ? 1 : 0;
          p->w_tx_bytes = k * nch * nb * rb; p->occ2 = pass == 0 ? 1 : 0;
          p->nb = nb; p->nch = nch; p->cs = cs; p->cb = cb; p->csp = nch * cb; p->k = k; p->cout = cout; p->two = 32 - k + 1;
        }
        return th;
      }
    }
  }
  return 0;
}

}  // namespace uegan

using namespace uegan;

extern "C" {

int uegan_conv2d_rowsum_supported(int32_t cout, int32_t cin_stored, int32_t k, int32_t dtype) {
  if ((dtype != UEGAN_F32 && dtype != UEGAN_F16) || cin_stored % 32 != 0 || !rowsum_shape_ok(cout, k)) return 0;
  const char* env = getenv("UEGAN_NO_ROWSUM");
  if (env && env[0] == '1') return 0;
  return rowsum_plan(cout, k, cin_stored, dtype, nullptr) != 0;
}

size_t uegan_packed_weight_rowsum_bytes(int32_t cout, int32_t cin_stored, int32_t k) {
  // (sized for either operand type: fp32 rows of cin_stored values >= fp16 rows padded to whole 64-channel chunks)
  return (size_t)rowsum_nb(cout, k) * k * ((cin_stored + 63) / 64 * 64) * sizeof(float);
}

static int pack_rowsum_impl(const float* w_oihw, void* w_packed, int cout, int cin_total, int cin_first, int cin,
                            int cin_stored, int k, int dtype, const float* wscale, void* stream) {
  UEGAN_CHECK(w_oihw && w_packed, "pack_conv_weight_rowsum: null pointer");
  UEGAN_CHECK(cin <= cin_stored && cin_first + cin <= cin_total && rowsum_shape_ok(cout, k),
              "pack_conv_weight_rowsum: unsupported shape (cout %d, k %d)", cout, k);
  UEGAN_CHECK(dtype == UEGAN_F32 || dtype == UEGAN_F16, "pack_conv_weight_rowsum: fp32 (tf32) or fp16 operands");
  const int nb = rowsum_nb(cout, k), cb = rowsum_cb(dtype);
  const int csp = (cin_stored + cb - 1) / cb * cb;
  const int total = nb * k * csp;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == UEGAN_F32)
    launch_pdl(pack_weight_rowsum_kernel<float>, (total + 255) / 256, 256, 0, st, w_oihw, static_cast<float*>(w_packed), cout,
                                                                         cin_total, cin_first, cin, csp, k, nb, wscale);
  else
    launch_pdl(pack_weight_rowsum_kernel<__half>, (total + 255) / 256, 256, 0, st, w_oihw, static_cast<__half*>(w_packed), cout,
                                                                          cin_total, cin_first, cin, csp, k, nb, wscale);
  UEGAN_CUDA(cudaGetLastError());
  return 0;
}

int uegan_pack_conv_weight_rowsum(const float* w_oihw, void* w_packed, int32_t cout, int32_t cin_total, int32_t cin_first,
                                  int32_t cin, int32_t cin_stored, int32_t k, void* stream) {
  return pack_rowsum_impl(w_oihw, w_packed, cout, cin_total, cin_first, cin, cin_stored, k, UEGAN_F32, nullptr, stream);
}

int uegan_pack_conv_weight_rowsum_scaled(const float* w_oihw, void* w_packed, int32_t cout, int32_t cin_total,
                                         int32_t cin_first, int32_t cin, int32_t cin_stored, int32_t k, int32_t dtype,
                                         const float* w_scale_dev, void* stream) {
  return pack_rowsum_impl(w_oihw, w_packed, cout, cin_total, cin_first, cin, cin_stored, k, dtype, w_scale_dev, stream);
}

int uegan_conv2d_fprop_rowsum(const uegan_conv_desc* desc, void* stream) {
  UEGAN_CHECK(desc != nullptr, "conv2d_fprop_rowsum: null desc");
  const uegan_conv_desc& d = *desc;
  const uegan_tensor& x = d.x;
  UEGAN_CHECK(x.data && d.w_packed && d.out_nchw, "conv2d_fprop_rowsum: null pointer (planar fp32 output only)");
  const bool f16 = x.dtype == UEGAN_F16;
  const int es = f16 ? 2 : 4;
  UEGAN_CHECK((x.dtype == UEGAN_F32 || f16) && x.c % 32 == 0 && d.stride == 1 && rowsum_shape_ok(d.cout, d.k) && !d.mul &&
                  !d.mask && !d.in_stats && d.y_mul <= 1,
              "conv2d_fprop_rowsum: unsupported convolution (cout %d, k %d, stride %d, c %d)", d.cout, d.k, d.stride, x.c);
  UEGAN_CHECK(d.pad <= x.halo, "conv2d_fprop_rowsum: pad %d exceeds input halo %d", d.pad, x.halo);
  RowsumParams p;
  memset(&p, 0, sizeof(p));
  UEGAN_CHECK(rowsum_plan(d.cout, d.k, x.c, x.dtype, &p) != 0,
              "conv2d_fprop_rowsum: weights do not fit in shared memory (c %d)", x.c);
  p.sx = x.scale; p.sw = d.w_scale;
  const int Ho = x.h + 2 * d.pad - d.k + 1, Wo = x.w + 2 * d.pad - d.k + 1;
  UEGAN_CHECK(Ho >= 1 && Wo >= 1, "conv2d_fprop_rowsum: empty output");
  p.Ho = Ho; p.Wo = Wo; p.act = d.act;
  p.tiles_w = (Wo + p.two - 1) / p.two;
  p.tiles_h = (Ho + p.th - 1) / p.th;
  p.total_tiles = p.tiles_w * p.tiles_h * x.n;
  p.patch_off = x.halo - d.pad;
  p.bias = d.bias; p.alpha = d.alpha;
  p.out_nchw = d.out_nchw; p.residual_nchw = d.residual_nchw; p.aux_nchw = d.aux_nchw;
  p.err_sink = error_sink_device();
  CUtensorMap tmA, tmB;
  {  // x, padded extent: {c, w, h, n}; box = one patch of 32 channels (coordinates beyond the tensor are zero-filled)
    // (fp16 tensors with 32 stored channels: the 64-channel box reads 32 real channels, the rest is zero-fill)
    const CUtensorMapDataType tdt = f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    const uint64_t pix = (uint64_t)x.c * es, row = (uint64_t)t_wp(x) * pix, img = (uint64_t)t_hp(x) * row;
    uint64_t dims[4] = {(uint64_t)x.c, (uint64_t)t_wp(x), (uint64_t)t_hp(x), (uint64_t)x.n};
    uint64_t strides[3] = {pix, row, img};
    const CUtensorMapSwizzle swz = p.row64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B;
    const uint32_t bc = p.row64 ? 32u : (uint32_t)p.cb;  // channels per operand row
    uint32_t box[4] = {bc, 32u, (uint32_t)p.ph, 1u};
    if (encode_tiled(&tmA, tdt, 4, x.data, dims, strides, box, swz)) return -1;
    const uint64_t ktot = (uint64_t)d.k * p.csp;
    uint64_t wdims[2] = {ktot, (uint64_t)p.nb};
    uint64_t wstrides[1] = {ktot * es};
    uint32_t wbox[2] = {bc, (uint32_t)p.nb};
    if (encode_tiled(&tmB, tdt, 2, const_cast<void*>(d.w_packed), wdims, wstrides, wbox, swz)) return -1;
  }
  static bool attr_set = false;
  if (!attr_set) {
    UEGAN_CUDA(cudaFuncSetAttribute(conv_rowsum_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024));
    UEGAN_CUDA(cudaFuncSetAttribute(conv_rowsum_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024));
    attr_set = true;
  }
  const int smem_bytes = p.w_total_bytes + p.num_stages * p.stage_bytes + 1024;
  const int max_ctas = p.occ2 ? 2 * num_sms() : num_sms();
  const int grid = p.total_tiles < max_ctas ? p.total_tiles : max_ctas;
  if (f16) launch_pdl(conv_rowsum_kernel<1>, grid, 384, smem_bytes, static_cast<cudaStream_t>(stream), tmA, tmB, p);
  else launch_pdl(conv_rowsum_kernel<0>, grid, 384, smem_bytes, static_cast<cudaStream_t>(stream), tmA, tmB, p);
  UEGAN_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
