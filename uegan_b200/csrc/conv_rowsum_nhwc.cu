// EXPERIMENTAL (written at the end of round 1 after the GPU budget was spent: compiles, NOT yet run on a B200; opt-in through
// UEGAN_ROWSUM_NHWC=1, never selected by default).
//
// Row-sum convolution (conv_rowsum.cu) generalised to Cout = 32 / 64 with an NHWC epilogue, for the k3 stride-1 layers whose
// MMAs are bound by the shared-memory fetch of their 128-row A slice at N = 32 / 64 (G's dec4, dec5.0; profiles/
// r1g_ncu_full_layers.md).  The three HORIZONTAL taps become GEMM columns, N = 3*Cout = 96 / 192:
//     D[p][(s, o)] = sum_r sum_c X[row(p) + r][col(p)][c] * w[o][c][r][s]
//     y[i][x][o]   = act(alpha * sum_s D[(i, x + s)][(s, o)] + bias[o]) (* mul[i][x][o])
// 24 MMAs of N = 96 per 240 output pixels instead of 72 of N = 32 per 128 (dec4): the A slice is fetched once per filter ROW.
// Patch = (th + 2) rows x 32 pixels x 32 channels per chunk (one 4-D TMA box), an M tile = 4 patch rows, tap r = the same patch
// through a descriptor advanced by r*4096 B -- exactly as in conv_rowsum_kernel.  Cout = 32: two M tiles per patch (th = 8),
// 128 TMEM columns each, two accumulator stages; Cout = 64: one M tile per patch (th = 4), 256 columns, two stages.
// Epilogue: lane = pixel of the patch row held by this warp's TMEM lane quarter; per 32-channel output group the three column
// groups (s = 0, 1, 2) are read, shifted with __shfl_down and summed; each lane stores its pixel's 32 channels (128 B).
#include "common.cuh"
#include "host_util.h"

namespace uegan {

constexpr int kRnStages = 4;
constexpr int kRnRowBytes = 32 * 128;

struct RowsumNhwcParams {
  int cout, nb;            // nb = 3 * cout
  int nch, cs;
  int th, msub, two;       // 8 / 2 (cout 32) or 4 / 1 (cout 64); 30 output columns per tile
  int tiles_w, tiles_h, total_tiles;
  int patch_off, ph;
  int stage_bytes, num_stages, stage_tx;
  int w_tile_bytes, w_total_bytes;
  int Wo, Ho, act;
  float* out;              // element (n = 0, y = 0, x = 0, c = y_c_off) of the interior
  long long out_pix, out_row, out_img;
  const float* mul;
  long long mul_pix, mul_row, mul_img;
  const float* bias;
  const float* alpha;
  unsigned int* err_sink;
};

__device__ __forceinline__ float rn_round_tf32(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return __uint_as_float(r);
}

// one warp, one M tile, one group of 32 output channels [og*32, og*32 + 32): lane = pixel column of the patch row
template <int COUT>
__device__ __forceinline__ void rowsum_nhwc_epilogue(const RowsumNhwcParams& p, uint32_t taddr, int og, long long o_off,
                                                     long long m_off, bool valid, float alpha) {
  float acc[32];
#pragma unroll
  for (int s = 0; s < 3; ++s) {
    uint32_t a[16], b[16];
    tmem_ld16(taddr + s * COUT + og * 32, a);
    tmem_ld16(taddr + s * COUT + og * 32 + 16, b);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float va = __uint_as_float(a[i]), vb = __uint_as_float(b[i]);
      if (s == 0) {
        acc[i] = va;
        acc[16 + i] = vb;
      } else {  // all 32 lanes shuffle (lanes >= 32 - s read their own value: masked by `valid`, lane < 30)
        acc[i] += __shfl_down_sync(0xffffffffu, va, s);
        acc[16 + i] += __shfl_down_sync(0xffffffffu, vb, s);
      }
    }
  }
  if (!valid) return;
  float* op = p.out + o_off + og * 32;
  const float* mp = p.mul ? p.mul + m_off + og * 32 : nullptr;
#pragma unroll
  for (int i = 0; i < 32; i += 4) {
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float x = acc[i + j] * alpha;
      if (p.bias) x += __ldg(p.bias + og * 32 + i + j);
      if (p.act == UEGAN_ACT_LRELU) x = fmaxf(x, 0.2f * x);
      else if (p.act == UEGAN_ACT_RELU) x = fmaxf(x, 0.f);
      v[j] = x;
    }
    if (mp) {
      const float4 m = __ldg(reinterpret_cast<const float4*>(mp + i));
      v[0] *= m.x; v[1] *= m.y; v[2] *= m.z; v[3] *= m.w;
    }
    *reinterpret_cast<float4*>(op + i) =
        make_float4(rn_round_tf32(v[0]), rn_round_tf32(v[1]), rn_round_tf32(v[2]), rn_round_tf32(v[3]));
  }
}

template <int COUT>
__global__ void __launch_bounds__(384, 1)
conv_rowsum_nhwc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                        const RowsumNhwcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t full_bar[kRnStages];
  __shared__ __align__(8) uint64_t empty_bar[kRnStages];
  __shared__ __align__(8) uint64_t tmem_full[2];
  __shared__ __align__(8) uint64_t tmem_empty[2];
  __shared__ __align__(8) uint64_t w_full;
  __shared__ uint32_t tmem_base_smem;
  constexpr int kSlot = COUT == 32 ? 128 : 256;  // TMEM columns per M tile (96 / 192 used)

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
  if (warp == 2) tmem_alloc(&tmem_base_smem, 512);  // 2 accumulator stages x 256 columns
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  const int tiles_per_img = p.tiles_w * p.tiles_h;

  if (warp == 0) {
    if (elect_one()) {
      // ===================== TMA producer =====================
      mbar_arrive_expect_tx(&w_full, (uint32_t)p.w_total_bytes);
      for (int wi = 0; wi < 3 * p.nch; ++wi) {  // weight tile (r, chunk): [nb rows x 128 B], resident
        const int r = wi / p.nch, c = wi % p.nch;
        tma_load_2d(&tmB, &w_full, smem + wi * p.w_tile_bytes, r * p.cs + c * 32, 0);
      }
      uint8_t* sa0 = smem + p.w_total_bytes;
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
        const int n = t / tiles_per_img, tt = t % tiles_per_img;
        const int wo0 = (tt % p.tiles_w) * p.two, ho0 = (tt / p.tiles_w) * p.th;
        for (int c = 0; c < p.nch; ++c) {
          mbar_wait(&empty_bar[stage], phase ^ 1, 0x1100 + stage, p.err_sink);
          mbar_arrive_expect_tx(&full_bar[stage], (uint32_t)p.stage_tx);
          tma_load_4d(&tmA, &full_bar[stage], sa0 + stage * p.stage_bytes, c * 32, wo0 + p.patch_off, ho0 + p.patch_off, n);
          if (++stage == p.num_stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      // ===================== MMA issuer =====================
      const uint32_t idesc = make_instr_desc(UMMA_TF32, 128, p.nb);
      const uint32_t w_addr = smem_u32(smem);
      const uint64_t db0 = make_smem_desc(w_addr, 16, 1024, UMMA_LAYOUT_SW128);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      mbar_wait(&w_full, 0, 0x1200, p.err_sink);
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1, 0x1210 + acc, p.err_sink);
        tcgen05_fence_after();
        for (int c = 0; c < p.nch; ++c) {
          mbar_wait(&full_bar[stage], phase, 0x1220 + stage, p.err_sink);
          tcgen05_fence_after();
          const uint64_t da0 = make_smem_desc(w_addr + p.w_total_bytes + stage * p.stage_bytes, 16, 1024, UMMA_LAYOUT_SW128);
          for (int j = 0; j < p.msub; ++j) {
            const uint32_t d_tmem = tmem_base + acc * 256 + j * kSlot;
            uint64_t da = desc_adv(da0, (uint32_t)(4 * j) * kRnRowBytes);
            uint64_t db = desc_adv(db0, (uint32_t)c * p.w_tile_bytes);
            const uint32_t b_step = (uint32_t)(p.nch * p.w_tile_bytes) >> 4;
#pragma unroll
            for (int r = 0; r < 3; ++r) {
              umma_ss<1>(d_tmem, da, db, idesc, (c | r) != 0 ? 1u : 0u);
              umma_ss<1>(d_tmem, da + 2, db + 2, idesc, 1u);
              umma_ss<1>(d_tmem, da + 4, db + 4, idesc, 1u);
              umma_ss<1>(d_tmem, da + 6, db + 6, idesc, 1u);
              da += kRnRowBytes >> 4;
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
    // ===================== epilogue =====================
    // q = patch row inside the M tile; Cout 32: half = M tile (two per patch); Cout 64: half = 32-channel output group
    const int q = warp & 3;
    const int half = (warp - 4) >> 2;
    const float alpha = p.alpha ? __ldg(p.alpha) : 1.0f;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
      const int n = t / tiles_per_img, tt = t % tiles_per_img;
      const int wo = (tt % p.tiles_w) * p.two + lane, ho0 = (tt / p.tiles_w) * p.th;
      const int j = COUT == 32 ? half : 0;
      const int og = COUT == 32 ? 0 : half;
      const int ho = ho0 + 4 * j + q;
      const bool valid = lane < p.two && wo < p.Wo && ho < p.Ho;
      const long long o_off = (long long)n * p.out_img + (long long)ho * p.out_row + (long long)wo * p.out_pix;
      const long long m_off = (long long)n * p.mul_img + (long long)ho * p.mul_row + (long long)wo * p.mul_pix;
      mbar_wait(&tmem_full[acc], acc_phase, 0x1300 + acc, p.err_sink);
      tcgen05_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * 256 + j * kSlot;
      rowsum_nhwc_epilogue<COUT>(p, taddr, og, o_off, m_off, valid, alpha);
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// out[(s, o)][r][c] (3*cout rows; 3*cs columns) = tf32(w[o][cin_first + c][r][s])
__global__ void pack_weight_rowsum_nhwc_kernel(const float* __restrict__ w, float* __restrict__ out, int cout, int cin_total,
                                               int cin_first, int cin, int cs) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 3 * cout * 3 * cs) return;
  const int c = i % cs, r = (i / cs) % 3, nrow = i / (cs * 3);
  const int s = nrow / cout, o = nrow % cout;
  float v = 0.f;
  if (c < cin) v = w[(((long long)o * cin_total + cin_first + c) * 3 + r) * 3 + s];
  out[i] = rn_round_tf32(v);
}

static int rowsum_nhwc_plan(int cout, int cs, RowsumNhwcParams* p) {
  if ((cout != 32 && cout != 64) || cs % 32 != 0) return 0;
  const int nb = 3 * cout, nch = cs / 32;
  const int th = cout == 32 ? 8 : 4;
  const long long w_total = 3LL * nch * nb * 128;
  const long long stage = (long long)(th + 2) * kRnRowBytes;
  const long long room = 200 * 1024 - w_total;
  if (room < 2 * stage) return 0;
  if (p) {
    p->cout = cout; p->nb = nb; p->nch = nch; p->cs = cs;
    p->th = th; p->msub = th / 4; p->two = 30; p->ph = th + 2;
    p->stage_bytes = (int)stage; p->stage_tx = (int)stage;
    p->num_stages = (int)(room / stage) < kRnStages ? (int)(room / stage) : kRnStages;
    p->w_tile_bytes = nb * 128; p->w_total_bytes = (int)w_total;
  }
  return th;
}

}  // namespace uegan

using namespace uegan;

extern "C" {

int uegan_conv2d_rowsum_nhwc_supported(int32_t cout, int32_t cin_stored, int32_t k, int32_t dtype) {
  const char* env = getenv("UEGAN_ROWSUM_NHWC");
  if (!(env && env[0] == '1')) return 0;  // experimental: opt-in only
  if (dtype != UEGAN_F32 || k != 3) return 0;
  return rowsum_nhwc_plan(cout, cin_stored, nullptr) != 0;
}

size_t uegan_packed_weight_rowsum_nhwc_bytes(int32_t cout, int32_t cin_stored) {
  return (size_t)3 * cout * 3 * cin_stored * sizeof(float);
}

int uegan_pack_conv_weight_rowsum_nhwc(const float* w_oihw, void* w_packed, int32_t cout, int32_t cin_total, int32_t cin_first,
                                       int32_t cin, int32_t cin_stored, void* stream) {
  UEGAN_CHECK(w_oihw && w_packed, "pack_conv_weight_rowsum_nhwc: null pointer");
  UEGAN_CHECK((cout == 32 || cout == 64) && cin <= cin_stored && cin_first + cin <= cin_total,
              "pack_conv_weight_rowsum_nhwc: unsupported shape (cout %d)", cout);
  const int total = 3 * cout * 3 * cin_stored;
  pack_weight_rowsum_nhwc_kernel<<<(total + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      w_oihw, static_cast<float*>(w_packed), cout, cin_total, cin_first, cin, cin_stored);
  UEGAN_CUDA(cudaGetLastError());
  return 0;
}

int uegan_conv2d_fprop_rowsum_nhwc(const uegan_conv_desc* desc, void* stream) {
  UEGAN_CHECK(desc != nullptr, "conv2d_fprop_rowsum_nhwc: null desc");
  const uegan_conv_desc& d = *desc;
  const uegan_tensor& x = d.x;
  const uegan_tensor& y = d.y;
  UEGAN_CHECK(x.data && y.data && d.w_packed, "conv2d_fprop_rowsum_nhwc: null pointer");
  UEGAN_CHECK(x.dtype == UEGAN_F32 && y.dtype == UEGAN_F32 && d.k == 3 && d.stride == 1 && !d.out_nchw && !d.mask &&
                  !d.in_stats && d.y_mul <= 1 && (d.act == UEGAN_ACT_NONE || d.act == UEGAN_ACT_LRELU || d.act == UEGAN_ACT_RELU),
              "conv2d_fprop_rowsum_nhwc: unsupported convolution");
  UEGAN_CHECK(d.pad <= x.halo, "conv2d_fprop_rowsum_nhwc: pad %d exceeds input halo %d", d.pad, x.halo);
  RowsumNhwcParams p;
  memset(&p, 0, sizeof(p));
  UEGAN_CHECK(rowsum_nhwc_plan(d.cout, x.c, &p) != 0, "conv2d_fprop_rowsum_nhwc: unsupported shape (cout %d, c %d)", d.cout, x.c);
  const int Ho = x.h + 2 * d.pad - 2, Wo = x.w + 2 * d.pad - 2;
  UEGAN_CHECK(y.n == x.n && y.h == Ho && y.w == Wo && d.y_c_off >= 0 && d.y_c_off % 4 == 0 && d.y_c_off + d.cout <= y.c,
              "conv2d_fprop_rowsum_nhwc: y is %dx%dx%dx%d, expected %dx%dx%d", y.n, y.h, y.w, y.c, x.n, Ho, Wo);
  p.Ho = Ho; p.Wo = Wo; p.act = d.act;
  p.tiles_w = (Wo + p.two - 1) / p.two;
  p.tiles_h = (Ho + p.th - 1) / p.th;
  p.total_tiles = p.tiles_w * p.tiles_h * x.n;
  p.patch_off = x.halo - d.pad;
  p.bias = d.bias; p.alpha = d.alpha;
  p.out_pix = y.c;
  p.out_row = t_wp(y) * y.c;
  p.out_img = t_hp(y) * p.out_row;
  p.out = static_cast<float*>(y.data) + (long long)y.halo * p.out_row + (long long)y.halo * p.out_pix + d.y_c_off;
  if (d.mul) {
    const uegan_tensor& mt = *d.mul;
    UEGAN_CHECK(mt.dtype == UEGAN_F32 && mt.n == y.n && mt.h == y.h && mt.w == y.w && mt.c >= d.cout && mt.c % 4 == 0,
                "conv2d_fprop_rowsum_nhwc: mul tensor mismatch");
    p.mul_pix = mt.c;
    p.mul_row = t_wp(mt) * mt.c;
    p.mul_img = t_hp(mt) * p.mul_row;
    p.mul = static_cast<const float*>(mt.data) + (long long)mt.halo * p.mul_row + (long long)mt.halo * p.mul_pix;
  }
  p.err_sink = error_sink_device();
  CUtensorMap tmA, tmB;
  {
    const uint64_t pix = (uint64_t)x.c * 4, row = (uint64_t)t_wp(x) * pix, img = (uint64_t)t_hp(x) * row;
    uint64_t dims[4] = {(uint64_t)x.c, (uint64_t)t_wp(x), (uint64_t)t_hp(x), (uint64_t)x.n};
    uint64_t strides[3] = {pix, row, img};
    uint32_t box[4] = {32u, 32u, (uint32_t)p.ph, 1u};
    if (encode_tiled(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, x.data, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B))
      return -1;
  }
  {
    const uint64_t ktot = 3ull * x.c;
    uint64_t dims[2] = {ktot, (uint64_t)p.nb};
    uint64_t strides[1] = {ktot * 4};
    uint32_t box[2] = {32u, (uint32_t)p.nb};
    if (encode_tiled(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(d.w_packed), dims, strides, box,
                     CU_TENSOR_MAP_SWIZZLE_128B))
      return -1;
  }
  static bool attr_set = false;
  if (!attr_set) {
    UEGAN_CUDA(cudaFuncSetAttribute(conv_rowsum_nhwc_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024));
    UEGAN_CUDA(cudaFuncSetAttribute(conv_rowsum_nhwc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024));
    attr_set = true;
  }
  const int smem_bytes = p.w_total_bytes + p.num_stages * p.stage_bytes + 1024;
  const int grid = p.total_tiles < num_sms() ? p.total_tiles : num_sms();
  if (d.cout == 32)
    conv_rowsum_nhwc_kernel<32><<<grid, 384, smem_bytes, static_cast<cudaStream_t>(stream)>>>(tmA, tmB, p);
  else
    conv_rowsum_nhwc_kernel<64><<<grid, 384, smem_bytes, static_cast<cudaStream_t>(stream)>>>(tmA, tmB, p);
  UEGAN_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
