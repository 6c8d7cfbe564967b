// Weight gradient of a convolution as a tcgen05 GEMM with BOTH operands MN-major (fp32 storage, kind::tf32):
//   dW[o][c][r][s] = alpha * sum_{n,ho,wo} dz[n,ho,wo,o] * xpad[n, ho*stride + r, wo*stride + s, c]
// (autograd of nn.Conv2d at models.py:83,94,162,174 -- `aten::convolution_backward`'s wgrad half).
// GEMM view per filter tap (r, s):  D[M = Cout tile (128)][N = Cin chunk (<= 256)] += A^T B, K = output pixels.
//   A = dz : NHWC, pixel-major  -> rows = pixels (K), 128-byte row = 32 channels (M)  == MN-major SWIZZLE_128B_ATOM_32B atoms
//   B = x  : NHWC window shifted by the tap, same layout with N = 32-channel groups
// One K stage = an 8x8 box of output pixels (64 rows): A = 4 TMA boxes {32 ch, 8, 8, 1} (M = 128), B = N/32 boxes;
// 8 MMAs (K = 8 pixels = one swizzle atom of rows) per stage.  Channels beyond the tensor are TMA zero-fill, so
// Cout < 128 costs no bandwidth.  Split-K over CTAs, fp32 atomics into the OIHW gradient (caller zeroes it).
// Small-Cin layers (RGB stored as 4 channels, enc1 / d1) use the sliding-window map of the fprop kernel: N = the 32
// (s, c) values of one filter row.  Tiny-Cout layers (G's last conv, D's prediction heads) swap the roles: M = 128
// consecutive (s, c) window elements of one filter row, N = the 32 stored dz channels -- k launches-worth of taps
// instead of k*k, and no MMA rows wasted on zero-filled output channels.
#include "common.cuh"
#include "host_util.h"

namespace uegan {

constexpr int kWgStageRows = 64;                      // pixels per K stage
constexpr int kWgBoxBytes = kWgStageRows * 128;       // one {128 bytes of channels x 64 px} box: 32 tf32 / 64 fp16 channels
// M = 128 channels: 4 boxes of 32 tf32 channels, 2 boxes of 64 fp16 channels
template <int kF16> struct WgT {
  static constexpr int CB = kF16 ? 64 : 32;           // channels per box (one 128-byte swizzle row)
  static constexpr int MBOX = 128 / CB;               // boxes of the M operand
  static constexpr int ABYTES = MBOX * kWgBoxBytes;
  static constexpr int KROWS = kF16 ? 16 : 8;         // pixels (K) per MMA
};

struct WgradParams {
  int tiles_w, tiles_h, nimg;      // 8x8 pixel tiles over (wo, ho, n)
  int total_ktiles, ksplit;
  int taps, k;                     // taps iterated by blockIdx.y: k*k, or k (filter rows) in window mode
  int n_boxes;                     // N / 32
  int num_stages, stage_bytes;
  int window_mode;                 // 1: small-Cin sliding-window B map
  int swap_mode;                   // 1: tiny Cout: M = 128 window elements (s, c) of one filter row, N = dz channels
  int tap_group, taps_total;       // non-swap: T taps share one MMA (N = T * per-tap columns), blockIdx.y = tap group
  int tap_boxes;                   // 32-column boxes per tap
  int cout, cin, cin_total, cin_first, x_c;
  int m_tiles, n_chunks;
  int a_bytes;                     // bytes of the M operand per stage (WgT::ABYTES)
  // N-operand boxes: 64 pixels x 128 B (b_cb = CB channels), or -- 32-channel fp16 x -- 64 pixels x 64 B (SWIZZLE_64B,
  // b_cb = 32): eight taps side by side fill N = 256 with real channels instead of four half-empty 64-channel boxes
  int b_cb, b_box_bytes;
  float* dw;                       // OIHW fp32
  float* partial;                  // deterministic mode: [k-split slice][dw_numel] partial sums, reduced in slice order
  long long dw_numel;
  const float* alpha;
  const float *sx, *sdz;           // per-tensor scales of x and dz (NULL = 1): the weight gradient is a true-scale value
  float scale;
  unsigned int* err_sink;
};

// split-K publication of one weight-gradient element: fp32 atomics (order = scheduling order, not reproducible), or a
// plain store into this k-slice's partial plane (wgrad_reduce_kernel then sums the planes in slice order: reproducible
// bit for bit, like the reference under cudnn.deterministic=True, utils.py:154)
__device__ __forceinline__ void wg_publish(float* dw, float* partial, long long dw_numel, long long idx, float v) {
  if (partial) partial[(long long)blockIdx.x * dw_numel + idx] = v;
  else atomicAdd(dw + idx, v);
}

// dw[o][cin_first + c][r][s] += sum_{slice} partial[slice][...] over the elements this launch owns.
// Block = E elements x L slice lanes (E * L = 256, L = 2^lanes_log2): lane l sums the slices l, l + L, ... in order (the
// loads of four slices are issued back to back), then the L lane sums are added in lane order -- a FIXED association
// (bit-reproducible) whose dependent-load chain is slices / (4 L) long: with up to 296 slices of a few thousand
// elements (1x1 convs) the one-thread-per-element version was pure latency (25-50 us per launch, 1.4 ms per step).
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(float* __restrict__ dw, const float* __restrict__ partial,
                                                           long long dw_numel, int slices, int cin_total, int cin_first,
                                                           int cin, int kk, int lanes_log2) {
  pdl_sync();
  __shared__ float sh[256];
  const int L = 1 << lanes_log2, E = 256 >> lanes_log2;
  const int e = threadIdx.x & (E - 1), l = threadIdx.x >> (8 - lanes_log2);
  // (32-bit index arithmetic: 64-bit divisions here cost more than the loads -- 130 us for D's d5, r2l)
  const unsigned i = blockIdx.x * (unsigned)E + (unsigned)e;
  float acc = 0.f;
  bool own = false;
  if (i < (unsigned)dw_numel) {
    const int c = (int)((i / (unsigned)kk) % (unsigned)cin_total);
    own = c >= cin_first && c < cin_first + cin;
    if (own) {
      const float* src = partial + i;
      int s = l;
      for (; s + 3 * L < slices; s += 4 * L) {
        const float a0 = __ldg(src + (long long)s * dw_numel), a1 = __ldg(src + (long long)(s + L) * dw_numel);
        const float a2 = __ldg(src + (long long)(s + 2 * L) * dw_numel), a3 = __ldg(src + (long long)(s + 3 * L) * dw_numel);
        acc = (((acc + a0) + a1) + a2) + a3;
      }
      for (; s < slices; s += L) acc += __ldg(src + (long long)s * dw_numel);
    }
  }
  sh[l * E + e] = acc;
  __syncthreads();
  if (l == 0 && own) {
    float t = sh[e];
    for (int k = 1; k < L; ++k) t += sh[k * E + e];
    dw[i] += t;
  }
}

template <int kF16>
__global__ void __launch_bounds__(256, 1)
conv_wgrad_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const WgradParams p) {
  constexpr int CB = WgT<kF16>::CB, MBOX = WgT<kF16>::MBOX, kWgABytes = WgT<kF16>::ABYTES, KROWS = WgT<kF16>::KROWS;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t full_bar[4];
  __shared__ __align__(8) uint64_t empty_bar[4];
  __shared__ __align__(8) uint64_t done_bar;
  __shared__ uint32_t tmem_base_smem;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // work decomposition: blockIdx.x = k-split slice, blockIdx.y = tap, blockIdx.z = m_tile * n_chunks + n_chunk
  const int tap = blockIdx.y;
  const int mt = blockIdx.z / p.n_chunks, nc = blockIdx.z % p.n_chunks;
  const int per = (p.total_ktiles + p.ksplit - 1) / p.ksplit;
  const int kt0 = blockIdx.x * per;
  const int kt1 = min(kt0 + per, p.total_ktiles);
  // swap mode: `tap` is the filter row.  Otherwise `tap` is a GROUP of up to tap_group taps (window mode: filter rows,
  // else (r, s) pairs) whose B boxes sit side by side in smem so that one MMA covers all of them (N = n_boxes * 32).
  const int r = tap;  // swap mode only
  const int t_first = tap * p.tap_group;
  const int t_count = p.swap_mode ? 1 : min(p.tap_group, p.taps_total - t_first);
  const int N = p.n_boxes * p.b_cb;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && elect_one()) {
    for (int i = 0; i < p.num_stages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    mbar_init(&done_bar, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(&tmem_base_smem, 256);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  pdl_sync();  // everything above is on-chip set-up: it overlaps the tail of the preceding kernel

  if (kt0 < kt1) {
    if (warp == 0) {
      if (elect_one()) {
        // The producer is ONE thread: everything that does not change from stage to stage (box coordinates of every
        // tap / channel group) is computed before the loop, and the tile counters advance by increments -- measured
        // r2j (ncu source page): with the coordinate arithmetic (divisions per box) inside the loop the MMA warp sat in
        // its full_bar wait for 80 % of its samples while this thread issued ~600 instructions per stage.
        constexpr int MAXB = 8;  // N-operand boxes per stage (256 columns of 32-channel boxes)
        int stage = 0;
        uint32_t phase = 0;
        const int nB = p.swap_mode ? 1 : t_count * p.tap_boxes;
        const uint32_t tx = kWgABytes + nB * p.b_box_bytes;
        int bc[MAXB], br[MAXB];
#pragma unroll
        for (int j = 0; j < MAXB; ++j) {
          const int t = j / p.tap_boxes, g = j - t * p.tap_boxes;
          const int tp = t_first + t;
          const int tr = p.window_mode ? tp : tp / p.k, ts = p.window_mode ? 0 : tp - (tp / p.k) * p.k;
          bc[j] = ts * p.x_c + nc * p.tap_boxes * p.b_cb + g * p.b_cb;
          br[j] = tr;
        }
        int tw_i = kt0 % p.tiles_w, th_i = (kt0 / p.tiles_w) % p.tiles_h, n = kt0 / (p.tiles_w * p.tiles_h);
        const int m0 = mt * 128;
        for (int kt = kt0; kt < kt1; ++kt) {
          const int wo0 = tw_i * 8, ho0 = th_i * 8;
          mbar_wait(&empty_bar[stage], phase ^ 1, 0x600 + stage, p.err_sink);
          uint8_t* sa = smem + stage * p.stage_bytes;
          uint8_t* sb = sa + kWgABytes;
          mbar_arrive_expect_tx(&full_bar[stage], tx);
          if (p.swap_mode) {
            // M operand = 128 consecutive window elements (s, c) of filter row r, N operand = the 32 stored dz channels
#pragma unroll
            for (int g = 0; g < MBOX; ++g)
              tma_load_5d(&tmB, &full_bar[stage], sa + g * kWgBoxBytes, m0 + g * CB, wo0, r, ho0, n);
            tma_load_4d(&tmA, &full_bar[stage], sb, 0, wo0, ho0, n);
          } else {
#pragma unroll
            for (int g = 0; g < MBOX; ++g)
              tma_load_4d(&tmA, &full_bar[stage], sa + g * kWgBoxBytes, m0 + g * CB, wo0, ho0, n);
            // B map = the fprop sliding-window map {window, wo, r, ho, n}: tap column s and channel offset are both
            // positions inside the window of k*C contiguous (s, c) values that starts at the pixel
#pragma unroll
            for (int j = 0; j < MAXB; ++j)
              if (j < nB) tma_load_5d(&tmB, &full_bar[stage], sb + j * p.b_box_bytes, bc[j], wo0, br[j], ho0, n);
          }
          if (++stage == p.num_stages) { stage = 0; phase ^= 1; }
          if (++tw_i == p.tiles_w) {
            tw_i = 0;
            if (++th_i == p.tiles_h) { th_i = 0; ++n; }
          }
        }
      }
    } else if (warp == 1) {
      if (elect_one()) {
        const uint32_t idesc = make_instr_desc(kF16 ? UMMA_F16 : UMMA_TF32, 128, N, /*a_mn_major=*/1, /*b_mn_major=*/1);
        int stage = 0;
        uint32_t phase = 0;
        uint32_t first = 0;
        for (int kt = kt0; kt < kt1; ++kt) {
          mbar_wait(&full_bar[stage], phase, 0x700 + stage, p.err_sink);
          tcgen05_fence_after();
          // MN-major tf32 operands exist only in the SWIZZLE_128B_BASE32B layout (32-byte chunks XORed with row & 3,
          // 4-row atoms; cutlass sm100_common.inl: "for mn-major tf32 operands, SW128_32B is the only available smem
          // layout"): LBO = stride between 32-channel groups (one TMA box), SBO = stride between 4-row groups.
          // MN-major fp16 operands: the plain SWIZZLE_128B layout, ((64 ch, n), (8 px, k)) : ((1, LBO), (128 B, SBO)):
          // 8-row atoms of 128-byte rows, SBO = 1024 bytes between them, LBO = one box between 64-channel groups.
          const uint64_t da0 = kF16 ? make_smem_desc(smem_u32(smem + stage * p.stage_bytes), kWgBoxBytes, 1024, UMMA_LAYOUT_SW128)
                                    : make_smem_desc(smem_u32(smem + stage * p.stage_bytes), kWgBoxBytes, 512, UMMA_LAYOUT_SW128_B32);
          const bool b64 = kF16 && p.b_box_bytes == kWgStageRows * 64;
          const uint64_t db0 = b64 ? make_smem_desc(smem_u32(smem + stage * p.stage_bytes) + kWgABytes, (uint32_t)p.b_box_bytes, 512,
                                                    UMMA_LAYOUT_SW64)
                                   : da0 + (kWgABytes >> 4);
          const uint32_t bstep = b64 ? KROWS * 4 : KROWS * 8;  // KROWS rows of 64 / 128 bytes, in 16-byte units
#pragma unroll
          for (int ks = 0; ks < kWgStageRows / KROWS; ++ks) {  // KROWS pixels (KROWS x 128 bytes of rows) per MMA
            umma_ss<kF16 ? 0 : 1>(tmem_base, da0 + ks * (KROWS * 8), db0 + ks * bstep, idesc, (first | ks) != 0 ? 1u : 0u);
          }
          first = 1;
          umma_commit(&empty_bar[stage]);
          if (++stage == p.num_stages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&done_bar);
      }
    } else if (warp >= 4) {
      // epilogue: D[o][j] -> atomicAdd into dW (OIHW)
      const int q = warp & 3;
      mbar_wait(&done_bar, 0, 0x800, p.err_sink);
      tcgen05_fence_after();
      const int o = mt * 128 + q * 32 + lane;
      const float sc = p.scale * (p.alpha ? __ldg(p.alpha) : 1.f) /
                       ((p.sx ? __ldg(p.sx) : 1.f) * (p.sdz ? __ldg(p.sdz) : 1.f));
      const int kk = p.k * p.k;
      for (int c0 = 0; c0 < N; c0 += 16) {
        uint32_t rr[16];
        tmem_ld16(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c0, rr);
        tmem_ld_wait();
        if (p.swap_mode) {
          const int j = o;  // window element (s, c) of filter row r
          const int ss = j / p.x_c, c = j % p.x_c;
          if (ss < p.k && c < p.cin) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int oc = c0 + i;
              if (oc < p.cout)
                wg_publish(p.dw, p.partial, p.dw_numel, ((long long)oc * p.cin_total + p.cin_first + c) * kk + r * p.k + ss,
                           __uint_as_float(rr[i]) * sc);
            }
          }
          continue;
        }
        if (o >= p.cout) continue;
        const int ncol = p.tap_boxes * p.b_cb;  // columns per tap
        const int tl = c0 / ncol;           // a 16-column chunk never straddles taps (ncol is a multiple of 32)
        if (tl >= t_count) continue;
        const int tp = t_first + tl;
        const int tr = p.window_mode ? tp : tp / p.k, ts = p.window_mode ? 0 : tp % p.k;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int j = (c0 + i) - tl * ncol;
          int c, ss;
          if (p.window_mode) { ss = j / p.x_c; c = j - ss * p.x_c; }  // window column = (s, c), x_c stored channels (RGB)
          else { ss = ts; c = nc * ncol + j; }
          if (c < p.cin && ss < p.k) {
            const float v = __uint_as_float(rr[i]) * sc;
            wg_publish(p.dw, p.partial, p.dw_numel, ((long long)o * p.cin_total + p.cin_first + c) * kk + tr * p.k + ss, v);
          }
        }
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 256);
}

// ------------------------------------------------------------------------------------------
// CUDA-core fallback for tiny spatial extents is not needed: TMA zero-fill covers partial tiles.
// ------------------------------------------------------------------------------------------

static int launch_wgrad_patch(const uegan_tensor* x, const uegan_tensor* dz, int cout, int cin, int cin_total,
                              int cin_first, int k, int pad, float* dw, const float* alpha, float scale, float* ws,
                              size_t ws_bytes, cudaStream_t stream);

// Deterministic mode (ws != NULL): clamp the k-split to what the workspace holds; after the GEMM launch, reduce the
// non-empty slices in order.
// A launch whose k-split ends up 1 needs neither: every dW element then has exactly ONE contributing thread, so the
// accumulation into dW itself (atomicAdd, *ws = NULL) is already reproducible.
static int wg_plan_split(int* ksplit, long long dw_numel, float** ws, size_t ws_bytes) {
  if (*ws) {
    UEGAN_CHECK(dw_numel < (1ll << 31), "conv2d_wgrad: weight tensor too large");
    const long long fit = (long long)(ws_bytes / (sizeof(float) * (size_t)dw_numel));
    UEGAN_CHECK(fit >= 1, "conv2d_wgrad: workspace of %zu bytes cannot hold one %lld-element partial", ws_bytes, dw_numel);
    if (*ksplit > fit) *ksplit = (int)fit;
    if (*ksplit == 1) *ws = nullptr;
  }
  return 0;
}
static thread_local int g_wgrad_launches = 0;  // kernels launched by the last weight-gradient call of this thread
static int wg_reduce(float* dw, const float* ws, long long dw_numel, int ksplit, int total_ktiles, int cin_total,
                     int cin_first, int cin, int k, cudaStream_t st) {
  g_wgrad_launches = ws ? 2 : 1;
  if (!ws) return 0;
  const int per = (total_ktiles + ksplit - 1) / ksplit;
  const int slices = (total_ktiles + per - 1) / per;  // slices beyond this one own no k-tile and write nothing
  // ~8 slices per thread: fewer, fuller blocks for the large weight tensors with a modest k-split (22 slices x 400 K
  // elements ran at 0.7 TB/s with 16 lanes), 32 lanes only for the 100+-slice launches
  int lanes_log2 = 0;
  while (lanes_log2 < 5 && (8 << lanes_log2) < slices) ++lanes_log2;
  const int E = 256 >> lanes_log2;
  launch_pdl(wgrad_reduce_kernel, (unsigned)((dw_numel + E - 1) / E), 256, 0, st, dw, ws, dw_numel, slices, cin_total, cin_first, cin,
                                                                         k * k, lanes_log2);
  UEGAN_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace uegan

using namespace uegan;

extern "C" int uegan_conv2d_wgrad(const uegan_tensor* x, const uegan_tensor* dz, int32_t cout, int32_t cin,
                                  int32_t cin_total, int32_t cin_first, int32_t k, int32_t stride, int32_t pad,
                                  float* dw_oihw, const float* alpha_dev, float scale, float* ws, size_t ws_bytes,
                                  void* stream) {
  UEGAN_CHECK(x && dz && x->data && dz->data && dw_oihw, "conv2d_wgrad: null pointer");
  UEGAN_CHECK(x->dtype == dz->dtype && (x->dtype == UEGAN_F32 || x->dtype == UEGAN_F16),
              "conv2d_wgrad: x and dz must both be fp32 (kind::tf32) or both fp16 (kind::f16)");
  const bool f16 = x->dtype == UEGAN_F16;
  const int es = f16 ? 2 : 4, CB = f16 ? 64 : 32;  // element size, channels per 128-byte box
  const int a_bytes = (128 / CB) * kWgBoxBytes;
  UEGAN_CHECK(stride == 1 || stride == 2, "conv2d_wgrad: stride %d", stride);
  UEGAN_CHECK(pad <= x->halo && k >= 1 && k <= 7, "conv2d_wgrad: bad pad/k");
  const int Ho = (x->h + 2 * pad - k) / stride + 1, Wo = (x->w + 2 * pad - k) / stride + 1;
  UEGAN_CHECK(dz->n == x->n && dz->h == Ho && dz->w == Wo, "conv2d_wgrad: dz is %dx%dx%d, expected %dx%dx%d", dz->n,
              dz->h, dz->w, x->n, Ho, Wo);
  UEGAN_CHECK(cout <= dz->c && cin <= x->c && cin_first + cin <= cin_total, "conv2d_wgrad: channel mismatch");
  // (channels beyond a tensor's extent are TMA zero-fill: any channel count whose pixel is a multiple of 16 bytes works;
  // whole 128-byte boxes are the efficient case)
  UEGAN_CHECK((dz->c * es) % 16 == 0 && (f16 || (dz->c * es) % 128 == 0),
              "conv2d_wgrad: unsupported dz channel count %d", dz->c);
  const bool window = (x->c * es == 16);  // RGB input stored as one 16-byte pixel
  if (stride == 1 && !window && k > 1) {
    const int took = launch_wgrad_patch(x, dz, cout, cin, cin_total, cin_first, k, pad, dw_oihw, alpha_dev, scale, ws,
                                        ws_bytes, static_cast<cudaStream_t>(stream));
    if (took != 0) return took < 0 ? -1 : 0;
  }
  UEGAN_CHECK(window || (f16 ? (x->c * es) % 16 == 0 : (x->c * 4) % 128 == 0),
              "conv2d_wgrad: unsupported x channel count %d", x->c);

  WgradParams p;
  memset(&p, 0, sizeof(p));
  p.tiles_w = (Wo + 7) / 8;
  p.tiles_h = (Ho + 7) / 8;
  p.nimg = x->n;
  p.total_ktiles = p.tiles_w * p.tiles_h * p.nimg;
  p.k = k;
  p.window_mode = window;
  p.taps = window ? k : k * k;
  p.cout = cout; p.cin = cin; p.cin_total = cin_total; p.cin_first = cin_first;
  p.m_tiles = (cout + 127) / 128;
  // role swap pays whenever the output-channel side would leave MMA rows empty (Cout <= 32 of M = 128)
  const bool swap = !window && cout <= 32 && dz->c == 32;
  p.swap_mode = swap;
  // 32-channel fp16 x (64-byte pixels): SWIZZLE_64B N-operand boxes
  const char* e64 = getenv("UEGAN_NO_WGRAD64");
  const bool b64 = f16 && !window && !swap && x->c == 32 && cin <= 32 && !(e64 && e64[0] == '1');
  int CBN = CB;  // channels per N-operand box
  int N;
  if (swap) {
    N = CB; p.n_chunks = 1;  // the (<= 32) dz channels: one box, zero-filled beyond the tensor
    p.taps = k;
    p.m_tiles = (k * x->c + 127) / 128;
    p.tap_group = 1; p.taps_total = k; p.tap_boxes = 1;
  } else if (window) {
    // RGB input: one "tap" = a filter row (the (s, c) window values of one box); rows side by side while N <= 256
    UEGAN_CHECK(k * x->c <= CB && k <= 8, "conv2d_wgrad: window mode needs k * stored channels <= %d", CB);
    p.taps_total = k;
    p.tap_boxes = 1;
    p.tap_group = k < 256 / CB ? k : 256 / CB;
    N = CB * p.tap_group; p.n_chunks = 1;
  } else {
    // several taps per MMA while the per-tap width leaves room in N <= 256 (A = dz is then read once per GROUP)
    if (b64) CBN = 32;
    const int CB = CBN;  // (shadows the 128-byte box width for this branch)
    const int cpad = (cin + CB - 1) / CB * CB;
    const int ncol = cpad < 256 ? cpad : 256;
    p.n_chunks = (cpad + ncol - 1) / ncol;
    p.tap_boxes = ncol / CB;
    p.taps_total = k * k;
    p.tap_group = 256 / ncol;
    if (p.tap_group > p.taps_total) p.tap_group = p.taps_total;
    N = p.tap_group * ncol;
  }
  p.b_cb = CBN;
  p.b_box_bytes = kWgStageRows * CBN * es;
  p.n_boxes = N / CBN;
  if (!swap) p.taps = (p.taps_total + p.tap_group - 1) / p.tap_group;
  p.a_bytes = a_bytes;
  p.stage_bytes = a_bytes + p.n_boxes * p.b_box_bytes;
  // Two CTAs per SM (each with half the smem ring and 256 of the 512 TMEM columns) when two stages still fit: the
  // launches are latency-bound (one TMA thread, one MMA thread, a short epilogue), and a second resident CTA overlaps them.
  const char* occ_env = getenv("UEGAN_WGRAD_OCC");
  const int occ = occ_env ? atoi(occ_env) : 2;
  const int groups = p.taps * p.m_tiles * p.n_chunks;
  // (only launches that are split into two waves' worth of CTAs anyway: a one-wave launch wants the deep ring; measured
  // r2z: D's d5 0.089 -> 0.118 ms with the halved ring, every two-wave launch 5 - 20 % faster)
  const int budget = (occ >= 2 && 2 * p.stage_bytes <= 100 * 1024 && 2 * groups < num_sms()) ? 100 * 1024 : 200 * 1024;
  p.num_stages = budget / p.stage_bytes;
  if (p.num_stages > 4) p.num_stages = 4;
  UEGAN_CHECK(p.num_stages >= 2, "conv2d_wgrad: stage too large");
  // two full waves of single-CTA SMs (rounding UP left a third, mostly empty wave); launches that fill half the SMs by
  // themselves are not split at all (no partial planes, no reduction pass)
  int ksplit = 2 * groups >= num_sms() ? 1 : (2 * num_sms()) / groups;
  // every CTA pays a fixed prologue + a 128 x N epilogue (and one more partial plane for the reduction): keep >= 16
  // K stages per CTA even when that leaves SMs idle (deep layers: few pixels, large dW)
  if (ksplit > p.total_ktiles / 16) ksplit = p.total_ktiles / 16;
  if (ksplit < 1) ksplit = 1;
  p.dw_numel = (long long)cout * cin_total * k * k;
  if (wg_plan_split(&ksplit, p.dw_numel, &ws, ws_bytes)) return -1;
  p.ksplit = ksplit;
  p.dw = dw_oihw;
  p.partial = ws;
  p.alpha = alpha_dev;
  p.scale = scale;
  p.sx = x->scale;
  p.sdz = dz->scale;
  p.err_sink = error_sink_device();

  p.x_c = x->c;
  CUtensorMap tmA, tmB;
  const CUtensorMapDataType tdt = f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  const CUtensorMapSwizzle tsw = f16 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
  {  // dz interior: {c, wo, ho, n}; coordinates beyond the interior / beyond Cout are zero-filled
    const uint64_t pix = (uint64_t)dz->c * es, row = (uint64_t)t_wp(*dz) * pix, img = (uint64_t)t_hp(*dz) * row;
    uint8_t* base = static_cast<uint8_t*>(dz->data) + (uint64_t)dz->halo * row + (uint64_t)dz->halo * pix;
    uint64_t dims[4] = {(uint64_t)dz->c, (uint64_t)Wo, (uint64_t)Ho, (uint64_t)dz->n};
    uint64_t strides[3] = {pix, row, img};
    uint32_t box[4] = {(uint32_t)CB, 8u, 8u, 1u};
    if (encode_tiled(&tmA, tdt, 4, base, dims, strides, box, tsw)) return -1;
  }
  {  // x: sliding-window map {window, wo, r, ho, n} (conv_fprop.cu)
    const uint64_t pix = (uint64_t)x->c * es, row = (uint64_t)t_wp(*x) * pix, img = (uint64_t)t_hp(*x) * row;
    uint8_t* base = static_cast<uint8_t*>(x->data) + (uint64_t)(x->halo - pad) * row + (uint64_t)(x->halo - pad) * pix;
    const uint64_t win = window ? (uint64_t)CB : (uint64_t)k * x->c;
    uint64_t dims[5] = {win, (uint64_t)Wo, (uint64_t)k, (uint64_t)Ho, (uint64_t)x->n};
    uint64_t strides[4] = {(uint64_t)stride * pix, row, (uint64_t)stride * row, img};
    uint32_t box[5] = {(uint32_t)CBN, 8u, 1u, 8u, 1u};
    if (encode_tiled(&tmB, tdt, 5, base, dims, strides, box, b64 ? CU_TENSOR_MAP_SWIZZLE_64B : tsw)) return -1;
  }
  static bool attr_set = false;
  if (!attr_set) {
    UEGAN_CUDA(cudaFuncSetAttribute(conv_wgrad_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024));
    UEGAN_CUDA(cudaFuncSetAttribute(conv_wgrad_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024));
    attr_set = true;
  }
  const int smem_bytes = p.num_stages * p.stage_bytes + 1024;
  dim3 grid((unsigned)p.ksplit, (unsigned)p.taps, (unsigned)(p.m_tiles * p.n_chunks));
  if (f16) launch_pdl(conv_wgrad_kernel<1>, grid, 256, smem_bytes, static_cast<cudaStream_t>(stream), tmA, tmB, p);
  else launch_pdl(conv_wgrad_kernel<0>, grid, 256, smem_bytes, static_cast<cudaStream_t>(stream), tmA, tmB, p);
  UEGAN_CUDA(cudaGetLastError());
  return wg_reduce(dw_oihw, ws, p.dw_numel, p.ksplit, p.total_ktiles, cin_total, cin_first, cin, k,
                   static_cast<cudaStream_t>(stream));
}

// ==========================================================================================================
// Patch ("Toeplitz descriptor") weight gradient for stride-1 convolutions of low-channel, high-resolution layers.
//   D[(r, s, c)][o] = sum_pix x[pix + (r, s)][c] * dz[pix][o]      (roles swapped: M = taps x channels, N = dz channels)
// One stage = an 8 x 8 tile of output pixels: per 32-channel chunk ONE halo'd patch of x ((8+k-1) rows x PW pixels, one
// 128-byte row per pixel) and the dz tile.  For filter row r, tile row ks and tap group g (4 taps), the A operand is
// read straight out of the patch by a descriptor whose M-group stride (LBO) is ONE PIXEL ROW (128 B): M-group j holds
// the 32 channels of tap s = 4g + j, K row i the pixel x0 + i, so element (j, i) is patch pixel (ks + r, i + 4g + j) --
// the sliding window is the descriptor, nothing is re-loaded per tap (k*k x less L2->SM traffic than one launch group
// per tap).  Accumulators: one 128 x N block of TMEM per (r, g, chunk); groups of them are split over blockIdx.y.
// ==========================================================================================================
namespace uegan {

struct WgradPatchParams {
  int tiles_w, tiles_h, nimg, total_ktiles, ksplit;
  int k, kgroups;              // kernel size, ceil(k / 4) tap groups
  int chunks;                  // x.c / 32
  int pw, ph;                  // patch width / height in pixels
  int patch_bytes, dz_bytes, stage_bytes, num_stages;
  int n_boxes;                 // N / 32
  int acc_total, acc_per_cta;  // (r, g, chunk) triples in total / per blockIdx.y slice
  int off;                     // halo - pad of x
  int cout, cin, cin_total, cin_first;
  // vertical ("h-stack") mode: dz is the horizontally unrolled gradient E[..][(s, o)] (uegan_dz_hstack) of a tiny-Cout
  // conv, so only the k VERTICAL taps remain: M-group j of tap group g = filter row 4g + j (LBO = one patch ROW), one
  // accumulator per (chunk, g), D[(r, c)][(s, o)] -> dW[o][c][r][s].  rk = taps enumerated by the accumulator index
  // besides the groups: k (horizontal mode: filter rows) or 1.
  int vert, rk, lbo_bytes;
  // columns of D: n_cols of them (a multiple of 16), column j = (j / col_c, j % col_c) = (horizontal tap, output channel);
  // col_rev: the tap index runs backwards (the stack is a sliding WINDOW over dz itself, see uegan_conv2d_wgrad_zwin)
  int n_cols, col_c, col_rev;
  int two_issuers;  // warps 1 and 3 both issue MMAs (alternate accumulators)
  int tmem_cols;    // TMEM columns the CTA allocates: 512, or 256 when its accumulators fit (two CTAs per SM then)
  // M-side operand rows: 128 bytes (cbm = CB channels, tpg = 128 / CB taps per M group), or -- 32-channel fp16 tensors,
  // vertical mode -- 64 bytes = one pixel (SWIZZLE_64B atoms of 32 channels: cbm = 32, tpg = 4 filter rows per M group:
  // no half-empty rows, half the MMAs)
  int cbm, tpg, a_row_bytes;
  float* dw;
  float* partial;
  long long dw_numel;
  const float* alpha;
  const float *sx, *sdz;  // per-tensor scales of x and dz (NULL = 1)
  float scale;
  unsigned int* err_sink;
};

// kF16: fp16 operands -- 64-channel chunks (one 128-byte row per pixel), M = 128 = TWO taps x 64 channels per group, K = 16
// pixels per MMA = two tile rows (K-atom stride SBO = one patch row for A, 1024 B for the dz tile).
template <int kF16>
__global__ void __launch_bounds__(256, 1)
conv_wgrad_patch_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmZ,
                        const WgradPatchParams p) {
  constexpr int CB = kF16 ? 64 : 32;       // channels per chunk
  constexpr int TPG = 128 / CB;            // taps per M = 128 group (4 tf32, 2 fp16)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t full_bar[8];
  __shared__ __align__(8) uint64_t empty_bar[8];
  __shared__ __align__(8) uint64_t done_bar;
  __shared__ uint32_t tmem_base_smem;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int per = (p.total_ktiles + p.ksplit - 1) / p.ksplit;
  const int kt0 = blockIdx.x * per, kt1 = min(kt0 + per, p.total_ktiles);
  const int acc0 = blockIdx.y * p.acc_per_cta, acc1 = min(acc0 + p.acc_per_cta, p.acc_total);
  const int N = p.n_cols;
  const int n_iss = (p.two_issuers && acc1 - acc0 >= 2) ? 2 : 1;
  // accumulator index a -> (chunk, r, g):  a = (chunk * k + r) * kgroups + g.  The chunks this CTA touches:
  const int ch_lo = acc0 / (p.rk * p.kgroups), ch_hi = (acc1 - 1) / (p.rk * p.kgroups);
  const int nch = ch_hi - ch_lo + 1;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmZ);
  }
  if (warp == 1 && elect_one()) {
    for (int i = 0; i < p.num_stages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], n_iss);
    }
    mbar_init(&done_bar, n_iss);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(&tmem_base_smem, p.tmem_cols);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  pdl_sync();  // everything above is on-chip set-up: it overlaps the tail of the preceding kernel

  if (kt0 < kt1) {
    if (warp == 0) {
      if (elect_one()) {
        int stage = 0;
        uint32_t phase = 0;
        const uint32_t tx = nch * p.ph * p.pw * p.a_row_bytes + p.n_boxes * 64 * 128;
        int tw_i = kt0 % p.tiles_w, th_i = (kt0 / p.tiles_w) % p.tiles_h, n_i = kt0 / (p.tiles_w * p.tiles_h);
        for (int kt = kt0; kt < kt1; ++kt) {
          const int wo0 = tw_i * 8, ho0 = th_i * 8, n = n_i;
          if (++tw_i == p.tiles_w) {
            tw_i = 0;
            if (++th_i == p.tiles_h) { th_i = 0; ++n_i; }
          }
          mbar_wait(&empty_bar[stage], phase ^ 1, 0xA00 + stage, p.err_sink);
          uint8_t* sp = smem + stage * p.stage_bytes;
          uint8_t* sz = sp + nch * p.patch_bytes;
          mbar_arrive_expect_tx(&full_bar[stage], tx);
          for (int c = 0; c < nch; ++c)
            tma_load_4d(&tmX, &full_bar[stage], sp + c * p.patch_bytes, (ch_lo + c) * p.cbm, wo0 + p.off, ho0 + p.off, n);
          for (int g = 0; g < p.n_boxes; ++g)
            tma_load_4d(&tmZ, &full_bar[stage], sz + g * 64 * 128, g * CB, wo0, ho0, n);
          if (++stage == p.num_stages) { stage = 0; phase ^= 1; }
        }
      }
    } else if (warp == 1 || warp == 3) {
      // Two MMA issuers (warp 3 joins when the CTA owns >= 2 accumulators): issuer i takes every second accumulator of
      // EVERY stage, so both walk the ring in lock step (each arrives once on empty_bar / done_bar) and no accumulator is
      // shared.  The one-thread issue loop (descriptor arithmetic + uniform-register moves per MMA), not the tensor pipe,
      // bounds these N <= 192 launches.
      const int iss = warp == 3 ? 1 : 0;
      if (iss < n_iss && elect_one()) {
        const uint32_t idesc = make_instr_desc(kF16 ? UMMA_F16 : UMMA_TF32, 128, N, 1, 1);
        int stage = 0;
        uint32_t phase = 0;
        uint32_t first = 0;
        for (int kt = kt0; kt < kt1; ++kt) {
          mbar_wait(&full_bar[stage], phase, 0xB00 + stage, p.err_sink);
          tcgen05_fence_after();
          const uint32_t sp = smem_u32(smem + stage * p.stage_bytes);
          const uint32_t sz = sp + nch * p.patch_bytes;
          // A: M-group stride (LBO) = one pixel row: the taps of a group are the same rows shifted by 0, 1, .. pixels.
          // tf32: 4-row K atoms inside one tile row (SBO 512); fp16: 8-row K atoms, the second one is the NEXT tile row
          // of the patch (SBO = patch row pitch; tcgen05 swizzles on absolute address bits, so any 128-byte multiple works)
          const uint64_t da0 = kF16 ? make_smem_desc(sp, (uint32_t)p.lbo_bytes, (uint32_t)(p.pw * p.a_row_bytes),
                                                     p.a_row_bytes == 64 ? UMMA_LAYOUT_SW64 : UMMA_LAYOUT_SW128)
                                    : make_smem_desc(sp, (uint32_t)p.lbo_bytes, 512, UMMA_LAYOUT_SW128_B32);
          const uint64_t db0 = kF16 ? make_smem_desc(sz, 64 * 128, 1024, UMMA_LAYOUT_SW128)
                                    : make_smem_desc(sz, 64 * 128, 512, UMMA_LAYOUT_SW128_B32);
          const uint32_t row_step = (uint32_t)(p.pw * p.a_row_bytes) >> 4;  // one patch row, in descriptor units of 16 bytes
          int g = acc0 % p.kgroups, r = (acc0 / p.kgroups) % p.rk, c = acc0 / (p.kgroups * p.rk) - ch_lo;
          uint32_t d_tmem = tmem_base;
          for (int a = acc0; a < acc1; ++a) {
            // first tap of the group: (row r, column TPG*g) of the patch, or (row TPG*g, column 0) in vertical mode
            const uint32_t tap0 = p.vert ? (uint32_t)(p.tpg * g * p.pw) : (uint32_t)(r * p.pw + p.tpg * g);
            uint64_t da = da0 + (uint32_t)(c * (p.patch_bytes >> 4)) + tap0 * (uint32_t)(p.a_row_bytes >> 4);
            constexpr int KSTEPS = kF16 ? 4 : 8;  // MMAs per 8x8 pixel tile: K = 8 (one tile row) or 16 (two tile rows)
            if (((a - acc0) & (n_iss - 1)) == iss) {
#pragma unroll
              for (int ks = 0; ks < KSTEPS; ++ks) {
                umma_ss<kF16 ? 0 : 1>(d_tmem, da, db0 + ks * (kF16 ? 128 : 64), idesc, (first | ks) != 0 ? 1u : 0u);
                da += kF16 ? 2 * row_step : row_step;
              }
            }
            d_tmem += N;
            if (++g == p.kgroups) { g = 0; if (++r == p.rk) { r = 0; ++c; } }
          }
          first = 1;
          umma_commit(&empty_bar[stage]);
          if (++stage == p.num_stages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&done_bar);
      }
    } else if (warp >= 4) {
      const int q = warp & 3;
      mbar_wait(&done_bar, 0, 0xC00, p.err_sink);
      tcgen05_fence_after();
      const int m = q * 32 + lane;        // row of D: tap j = m / CB of the group, channel m % CB of the chunk
      const float sc = p.scale * (p.alpha ? __ldg(p.alpha) : 1.f) /
                       ((p.sx ? __ldg(p.sx) : 1.f) * (p.sdz ? __ldg(p.sdz) : 1.f));
      const int kk = p.k * p.k;
      for (int a = acc0; a < acc1; ++a) {
        const int g = a % p.kgroups, r = (a / p.kgroups) % p.rk, chunk = a / (p.kgroups * p.rk);
        const int s_ = p.tpg * g + m / p.cbm, c = chunk * p.cbm + (m % p.cbm);  // vertical mode: s_ is the filter ROW
        for (int c0 = 0; c0 < N; c0 += 16) {
          uint32_t rr[16];
          tmem_ld16(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + (a - acc0) * N + c0, rr);
          tmem_ld_wait();
          if (s_ >= p.k || c >= p.cin) continue;
          if (p.vert) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int j = c0 + i;  // column (s, o) of the stacked gradient
              const int sj = j / p.col_c, o = j - sj * p.col_c;
              if (sj < p.k && o < p.cout) {
                const int ss = p.col_rev ? p.k - 1 - sj : sj;
                wg_publish(p.dw, p.partial, p.dw_numel, ((long long)o * p.cin_total + p.cin_first + c) * kk + s_ * p.k + ss,
                           __uint_as_float(rr[i]) * sc);
              }
            }
            continue;
          }
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int o = c0 + i;
            if (o < p.cout)
              wg_publish(p.dw, p.partial, p.dw_numel, ((long long)o * p.cin_total + p.cin_first + c) * kk + r * p.k + s_,
                         __uint_as_float(rr[i]) * sc);
          }
        }
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, p.tmem_cols);
}

// returns 1 if the patch kernel took the job, 0 if the caller should use the generic kernel, -1 on error
template <int kF16>
static int wgrad_patch_set_attr() {
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(conv_wgrad_patch_kernel<kF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024) !=
        cudaSuccess)
      return set_error("conv2d_wgrad: cudaFuncSetAttribute failed");
    attr_set = true;
  }
  return 0;
}

static int launch_wgrad_patch(const uegan_tensor* x, const uegan_tensor* dz, int cout, int cin, int cin_total,
                              int cin_first, int k, int pad, float* dw, const float* alpha, float scale, float* ws,
                              size_t ws_bytes, cudaStream_t stream) {
  const char* env = getenv("UEGAN_NO_WGRAD_PATCH");
  if (env && env[0] == '1') return 0;
  const bool f16 = x->dtype == UEGAN_F16;
  const int es = f16 ? 2 : 4, CB = f16 ? 64 : 32, TPG = 128 / CB;
  if (x->c % 32 != 0 || dz->c % 32 != 0 || cout > 64 || k > 8) return 0;
  const int Ho = x->h + 2 * pad - k + 1, Wo = x->w + 2 * pad - k + 1;
  if (Ho < 8 || Wo < 8) return 0;
  WgradPatchParams p;
  memset(&p, 0, sizeof(p));
  const int N = (cout + CB - 1) / CB * CB;
  p.n_boxes = N / CB;
  p.k = k;
  p.kgroups = (k + TPG - 1) / TPG;
  p.chunks = (cin + CB - 1) / CB;
  p.pw = 8 + TPG * p.kgroups - 1;
  p.ph = 8 + k - 1;
  p.patch_bytes = (p.ph * p.pw * 128 + 1023) / 1024 * 1024;
  p.dz_bytes = p.n_boxes * 64 * 128;
  p.vert = 0; p.rk = k; p.lbo_bytes = 128;
  p.n_cols = N; p.col_c = cout; p.col_rev = 0;
  p.acc_total = p.chunks * k * p.kgroups;
  p.acc_per_cta = 512 / N;
  // a CTA's accumulators should span as few channel chunks as possible: align the slice to whole chunks when it can
  const int per_chunk = k * p.kgroups;
  if (p.acc_per_cta >= per_chunk) p.acc_per_cta = (p.acc_per_cta / per_chunk) * per_chunk;
  if (p.acc_per_cta > p.acc_total) p.acc_per_cta = p.acc_total;
  const int slices = (p.acc_total + p.acc_per_cta - 1) / p.acc_per_cta;
  const int max_nch = p.acc_per_cta >= per_chunk ? p.acc_per_cta / per_chunk : 2;
  p.stage_bytes = max_nch * p.patch_bytes + p.dz_bytes;
  p.num_stages = (200 * 1024) / p.stage_bytes;
  if (p.num_stages > 4) p.num_stages = 4;
  if (p.num_stages < 2) return 0;
  p.tiles_w = (Wo + 7) / 8;
  p.tiles_h = (Ho + 7) / 8;
  p.nimg = x->n;
  p.total_ktiles = p.tiles_w * p.tiles_h * p.nimg;
  int ksplit = (2 * num_sms()) / slices;
  if (ksplit > p.total_ktiles) ksplit = p.total_ktiles;
  if (ksplit < 1) ksplit = 1;
  p.dw_numel = (long long)cout * cin_total * k * k;
  if (wg_plan_split(&ksplit, p.dw_numel, &ws, ws_bytes)) return -1;
  p.ksplit = ksplit;
  p.off = x->halo - pad;
  p.cout = cout; p.cin = cin; p.cin_total = cin_total; p.cin_first = cin_first;
  p.dw = dw; p.partial = ws; p.alpha = alpha; p.scale = scale;
  p.sx = x->scale; p.sdz = dz->scale;
  p.err_sink = error_sink_device();
  { const char* e2 = getenv("UEGAN_WGRAD_ISSUERS"); p.two_issuers = !(e2 && e2[0] == '1'); }
  if (p.tmem_cols == 0) p.tmem_cols = 512;
  if (p.cbm == 0) { p.cbm = CB; p.tpg = TPG; p.a_row_bytes = 128; }
  const CUtensorMapDataType tdt = f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  const CUtensorMapSwizzle tsw = f16 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
  CUtensorMap tmX, tmZ;
  {  // x, padded extent: {c, w, h, n}
    const uint64_t pix = (uint64_t)x->c * es, row = (uint64_t)t_wp(*x) * pix, img = (uint64_t)t_hp(*x) * row;
    uint64_t dims[4] = {(uint64_t)x->c, (uint64_t)t_wp(*x), (uint64_t)t_hp(*x), (uint64_t)x->n};
    uint64_t strides[3] = {pix, row, img};
    uint32_t box[4] = {(uint32_t)CB, (uint32_t)p.pw, (uint32_t)p.ph, 1u};
    if (encode_tiled(&tmX, tdt, 4, x->data, dims, strides, box, tsw)) return -1;
  }
  {  // dz interior: {c, wo, ho, n}
    const uint64_t pix = (uint64_t)dz->c * es, row = (uint64_t)t_wp(*dz) * pix, img = (uint64_t)t_hp(*dz) * row;
    uint8_t* base = static_cast<uint8_t*>(dz->data) + (uint64_t)dz->halo * row + (uint64_t)dz->halo * pix;
    uint64_t dims[4] = {(uint64_t)dz->c, (uint64_t)Wo, (uint64_t)Ho, (uint64_t)dz->n};
    uint64_t strides[3] = {pix, row, img};
    uint32_t box[4] = {(uint32_t)CB, 8u, 8u, 1u};
    if (encode_tiled(&tmZ, tdt, 4, base, dims, strides, box, tsw)) return -1;
  }
  const int smem_bytes = p.num_stages * p.stage_bytes + 1024;
  dim3 grid((unsigned)p.ksplit, (unsigned)slices, 1u);
  if (f16) {
    if (wgrad_patch_set_attr<1>()) return -1;
    launch_pdl(conv_wgrad_patch_kernel<1>, grid, 256, smem_bytes, stream, tmX, tmZ, p);
  } else {
    if (wgrad_patch_set_attr<0>()) return -1;
    launch_pdl(conv_wgrad_patch_kernel<0>, grid, 256, smem_bytes, stream, tmX, tmZ, p);
  }
  if (cudaGetLastError() != cudaSuccess) return set_error("conv2d_wgrad: patch kernel launch failed");
  if (wg_reduce(dw, ws, p.dw_numel, p.ksplit, p.total_ktiles, cin_total, cin_first, cin, k, stream)) return -1;
  return 1;
}

}  // namespace uegan

// Weight gradient of a stride-1 conv with a tiny output-channel count from the horizontally unrolled gradient
// E = uegan_dz_hstack(dz) (32 stored channels = (s, o) pairs): the patch kernel in vertical mode.
extern "C" int uegan_conv2d_wgrad_hstack(const uegan_tensor* x, const uegan_tensor* e, int32_t cout, int32_t cin,
                                         int32_t cin_total, int32_t cin_first, int32_t k, int32_t pad, float* dw_oihw,
                                         const float* alpha_dev, float scale, float* ws, size_t ws_bytes, void* stream) {
  UEGAN_CHECK(x && e && x->data && e->data && dw_oihw, "conv2d_wgrad_hstack: null pointer");
  UEGAN_CHECK(x->dtype == e->dtype && (x->dtype == UEGAN_F32 || x->dtype == UEGAN_F16) && x->c % 32 == 0 && e->c == 32 &&
                  e->halo == 0,
              "conv2d_wgrad_hstack: fp32 or fp16 tensors, x.c %% 32 == 0, 32-channel stack without halo");
  const bool f16 = x->dtype == UEGAN_F16;
  const int es = f16 ? 2 : 4, CB = f16 ? 64 : 32, TPG = 128 / CB;
  UEGAN_CHECK(k >= 1 && k <= 7 && (k & 1) && pad == (k - 1) / 2 && pad <= x->halo && k * cout <= 32,
              "conv2d_wgrad_hstack: unsupported k %d / pad %d / cout %d", k, pad, cout);
  UEGAN_CHECK(e->n == x->n && e->h == x->h && e->w == x->w + k - 1, "conv2d_wgrad_hstack: stack is %dx%dx%d, expected %dx%dx%d",
              e->n, e->h, e->w, x->n, x->h, x->w + k - 1);
  UEGAN_CHECK(cin <= x->c && cin_first + cin <= cin_total, "conv2d_wgrad_hstack: channel mismatch");
  WgradPatchParams p;
  memset(&p, 0, sizeof(p));
  p.n_boxes = 1;  // the 32 (s, o) columns of the stack: one box (fp16: zero-filled beyond the 32 stored channels)
  p.k = k;
  p.kgroups = (k + TPG - 1) / TPG;
  p.chunks = (cin + CB - 1) / CB;
  p.vert = 1; p.rk = 1;
  p.n_cols = CB; p.col_c = cout; p.col_rev = 0;
  p.pw = 8;
  p.ph = 8 + TPG * p.kgroups - 1;     // rows 0 .. 7 + (TPG*kgroups - 1): M-group j of the last group stays inside
  p.lbo_bytes = p.pw * 128;           // M-group stride = one patch row
  p.patch_bytes = p.ph * p.pw * 128;  // a multiple of 1024
  p.dz_bytes = 64 * 128;
  p.acc_total = p.chunks * p.kgroups;
  const int N = CB;
  // channel chunks per CTA: at least 3 stages of (chunks * patch + stack tile) in 200 KB, at most 512 / N accumulators
  int nch_max = ((200 * 1024) / 3 - p.dz_bytes) / p.patch_bytes;
  if (nch_max > (512 / N) / p.kgroups) nch_max = (512 / N) / p.kgroups;
  if (nch_max < 1) nch_max = 1;
  const int slices = (p.chunks + nch_max - 1) / nch_max;
  const int nch = (p.chunks + slices - 1) / slices;
  p.acc_per_cta = nch * p.kgroups;
  p.stage_bytes = nch * p.patch_bytes + p.dz_bytes;
  p.num_stages = (200 * 1024) / p.stage_bytes;
  if (p.num_stages > 4) p.num_stages = 4;
  UEGAN_CHECK(p.num_stages >= 2, "conv2d_wgrad_hstack: stage too large");
  p.tiles_w = (e->w + 7) / 8;
  p.tiles_h = (e->h + 7) / 8;
  p.nimg = x->n;
  p.total_ktiles = p.tiles_w * p.tiles_h * p.nimg;
  int ksplit = (2 * num_sms()) / slices;
  if (ksplit > p.total_ktiles) ksplit = p.total_ktiles;
  if (ksplit < 1) ksplit = 1;
  p.dw_numel = (long long)cout * cin_total * k * k;
  if (wg_plan_split(&ksplit, p.dw_numel, &ws, ws_bytes)) return -1;
  p.ksplit = ksplit;
  p.off = x->halo - pad;
  p.cout = cout; p.cin = cin; p.cin_total = cin_total; p.cin_first = cin_first;
  p.dw = dw_oihw; p.partial = ws; p.alpha = alpha_dev; p.scale = scale;
  p.sx = x->scale; p.sdz = e->scale;
  p.err_sink = error_sink_device();
  { const char* e2 = getenv("UEGAN_WGRAD_ISSUERS"); p.two_issuers = !(e2 && e2[0] == '1'); }
  if (p.tmem_cols == 0) p.tmem_cols = 512;
  if (p.cbm == 0) { p.cbm = CB; p.tpg = TPG; p.a_row_bytes = 128; }
  const CUtensorMapDataType tdt = f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  const CUtensorMapSwizzle tsw = f16 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
  CUtensorMap tmX, tmZ;
  {  // x, padded extent: {c, w, h, n}
    const uint64_t pix = (uint64_t)x->c * es, row = (uint64_t)t_wp(*x) * pix, img = (uint64_t)t_hp(*x) * row;
    uint64_t dims[4] = {(uint64_t)x->c, (uint64_t)t_wp(*x), (uint64_t)t_hp(*x), (uint64_t)x->n};
    uint64_t strides[3] = {pix, row, img};
    uint32_t box[4] = {(uint32_t)CB, (uint32_t)p.pw, (uint32_t)p.ph, 1u};
    if (encode_tiled(&tmX, tdt, 4, x->data, dims, strides, box, tsw)) return -1;
  }
  {  // the stack: {32, w + k - 1, h, n}
    const uint64_t pix = 32 * es, row = (uint64_t)e->w * pix, img = (uint64_t)e->h * row;
    uint64_t dims[4] = {32u, (uint64_t)e->w, (uint64_t)e->h, (uint64_t)e->n};
    uint64_t strides[3] = {pix, row, img};
    uint32_t box[4] = {(uint32_t)CB, 8u, 8u, 1u};
    if (encode_tiled(&tmZ, tdt, 4, e->data, dims, strides, box, tsw)) return -1;
  }
  const int smem_bytes = p.num_stages * p.stage_bytes + 1024;
  dim3 grid((unsigned)p.ksplit, (unsigned)slices, 1u);
  if (f16) {
    if (wgrad_patch_set_attr<1>()) return -1;
    launch_pdl(conv_wgrad_patch_kernel<1>, grid, 256, smem_bytes, static_cast<cudaStream_t>(stream), tmX, tmZ, p);
  } else {
    if (wgrad_patch_set_attr<0>()) return -1;
    launch_pdl(conv_wgrad_patch_kernel<0>, grid, 256, smem_bytes, static_cast<cudaStream_t>(stream), tmX, tmZ, p);
  }
  UEGAN_CUDA(cudaGetLastError());
  return wg_reduce(dw_oihw, ws, p.dw_numel, p.ksplit, p.total_ktiles, cin_total, cin_first, cin, k,
                   static_cast<cudaStream_t>(stream));
}

// Weight gradient of a stride-1 conv from a SLIDING WINDOW over dz itself (no materialised stack): with a zero halo of
// >= k - 1 pixels around dz, the k * Cs contiguous values that start at pixel q - (k - 1) of a dz row ARE the stack row
//   E[q][(s', o)] = dz[q - (k - 1) + s'][o] = dz[q - s][o],  s = k - 1 - s',  q in [0, W + k - 1)
// (Cs = stored channels of dz), so one overlapping-stride TMA map delivers the N operand of the vertical patch mode:
//   D[(r, c)][(s', o)] = sum_{y, q} xpad[y + r][q][c] * E[y][q][(s', o)] = dW[o][c][r][k - 1 - s'].
// Only the k VERTICAL taps are enumerated on the M side (k x fewer MMAs than one accumulator per (r, s), each k x wider:
// Cout = 32, k = 3 runs N = 96 instead of 6 MMAs of N = 64 per K step).  Needs k * Cs <= 256 columns.
extern "C" int uegan_conv2d_wgrad_zwin_supported(int32_t cout, int32_t dz_c, int32_t dz_halo, int32_t x_c, int32_t k,
                                                 int32_t stride, int32_t dtype) {
  const int ncols = (k * dz_c + 15) / 16 * 16;
  return dtype == UEGAN_F16 && stride == 1 && k >= 3 && k <= 7 && (k & 1) && dz_halo >= k - 1 && dz_c % 8 == 0 &&
         cout <= dz_c && ncols <= 256 && x_c % 32 == 0;
}

extern "C" int uegan_conv2d_wgrad_zwin(const uegan_tensor* x, const uegan_tensor* dz, int32_t cout, int32_t cin,
                                       int32_t cin_total, int32_t cin_first, int32_t k, int32_t pad, float* dw_oihw,
                                       const float* alpha_dev, float scale, float* ws, size_t ws_bytes, void* stream) {
  UEGAN_CHECK(x && dz && x->data && dz->data && dw_oihw, "conv2d_wgrad_zwin: null pointer");
  UEGAN_CHECK(x->dtype == UEGAN_F16 && dz->dtype == UEGAN_F16, "conv2d_wgrad_zwin: fp16 tensors only");
  UEGAN_CHECK(uegan_conv2d_wgrad_zwin_supported(cout, dz->c, dz->halo, x->c, k, 1, x->dtype),
              "conv2d_wgrad_zwin: unsupported shape (cout %d, dz.c %d, dz.halo %d, x.c %d, k %d)", cout, dz->c, dz->halo, x->c, k);
  UEGAN_CHECK(pad == (k - 1) / 2 && pad <= x->halo, "conv2d_wgrad_zwin: pad %d (k %d, x.halo %d)", pad, k, x->halo);
  UEGAN_CHECK(dz->n == x->n && dz->h == x->h && dz->w == x->w, "conv2d_wgrad_zwin: dz is %dx%dx%d, expected %dx%dx%d", dz->n,
              dz->h, dz->w, x->n, x->h, x->w);
  UEGAN_CHECK(cin <= x->c && cin_first + cin <= cin_total, "conv2d_wgrad_zwin: channel mismatch");
  constexpr int es = 2, CB = 64;
  WgradPatchParams p;
  memset(&p, 0, sizeof(p));
  // 32-channel x: a pixel is 64 bytes -- SWIZZLE_64B rows, four filter rows per M = 128 group instead of two half-empty ones
  const char* e64 = getenv("UEGAN_NO_ZWIN64");
  const bool row64 = x->c == 32 && !(e64 && e64[0] == '1');
  const int CBM = row64 ? 32 : 64, TPG = 128 / CBM, arow = row64 ? 64 : 128;
  p.cbm = CBM; p.tpg = TPG; p.a_row_bytes = arow;
  p.n_cols = (k * dz->c + 15) / 16 * 16;
  p.n_boxes = (p.n_cols + CB - 1) / CB;
  p.col_c = dz->c; p.col_rev = 1;
  p.k = k;
  p.kgroups = (k + TPG - 1) / TPG;
  p.chunks = (cin + CBM - 1) / CBM;
  p.vert = 1; p.rk = 1;
  p.pw = 8;
  p.ph = 8 + TPG * p.kgroups - 1;
  p.lbo_bytes = p.pw * arow;
  p.patch_bytes = (p.ph * p.pw * arow + 1023) / 1024 * 1024;
  p.dz_bytes = p.n_boxes * 64 * 128;
  p.acc_total = p.chunks * p.kgroups;
  const int N = p.n_cols;
  int nch_max = ((200 * 1024) / 3 - p.dz_bytes) / p.patch_bytes;
  if (nch_max > (512 / N) / p.kgroups) nch_max = (512 / N) / p.kgroups;
  UEGAN_CHECK(nch_max >= 1, "conv2d_wgrad_zwin: %d columns x %d tap groups do not fit the accumulator memory", N, p.kgroups);
  const int slices = (p.chunks + nch_max - 1) / nch_max;
  const int nch = (p.chunks + slices - 1) / slices;
  p.acc_per_cta = nch * p.kgroups;
  p.stage_bytes = nch * p.patch_bytes + p.dz_bytes;
  {
    // several CTAs per SM (the launch is latency-bound: one TMA thread, one or two MMA threads, a short epilogue): the
    // largest count whose accumulators share the 512 TMEM columns and whose rings (>= 3 stages for two CTAs, >= 2 beyond)
    // share the shared memory
    const char* occ_env = getenv("UEGAN_WGRAD_OCC");
    const int occ_max = occ_env ? atoi(occ_env) : 2;
    int cols = 32;
    while (cols < p.acc_per_cta * N) cols <<= 1;
    int occ = 1;
    for (int o = occ_max < 4 ? occ_max : 4; o >= 2; --o) {
      const int budget = (o == 2 ? 100 : (o == 3 ? 66 : 49)) * 1024;
      if (cols * o <= 512 && (o == 2 ? 3 : 2) * p.stage_bytes <= budget) { occ = o; break; }
    }
    p.tmem_cols = occ == 1 ? 512 : cols < 512 / occ ? (occ == 3 ? 128 : 512 / occ) : 512 / occ;
    if (occ == 3 && cols > 128) p.tmem_cols = cols;  // (three CTAs: 3 x cols <= 512 already checked)
    const int budget = occ == 1 ? 200 * 1024 : (occ == 2 ? 100 : (occ == 3 ? 66 : 49)) * 1024;
    p.num_stages = budget / p.stage_bytes;
  }
  if (p.num_stages > 6) p.num_stages = 6;
  UEGAN_CHECK(p.num_stages >= 2, "conv2d_wgrad_zwin: stage too large");
  const int We = dz->w + k - 1;
  p.tiles_w = (We + 7) / 8;
  p.tiles_h = (dz->h + 7) / 8;
  p.nimg = x->n;
  p.total_ktiles = p.tiles_w * p.tiles_h * p.nimg;
  int ksplit = 2 * slices >= num_sms() ? 1 : (2 * num_sms()) / slices;
  if (ksplit > p.total_ktiles / 16) ksplit = p.total_ktiles / 16;
  if (ksplit < 1) ksplit = 1;
  p.dw_numel = (long long)cout * cin_total * k * k;
  if (wg_plan_split(&ksplit, p.dw_numel, &ws, ws_bytes)) return -1;
  p.ksplit = ksplit;
  p.off = x->halo - pad;
  p.cout = cout; p.cin = cin; p.cin_total = cin_total; p.cin_first = cin_first;
  p.dw = dw_oihw; p.partial = ws; p.alpha = alpha_dev; p.scale = scale;
  p.sx = x->scale; p.sdz = dz->scale;
  p.err_sink = error_sink_device();
  { const char* e2 = getenv("UEGAN_WGRAD_ISSUERS"); p.two_issuers = !(e2 && e2[0] == '1'); }
  if (p.tmem_cols == 0) p.tmem_cols = 512;
  if (p.cbm == 0) { p.cbm = CB; p.tpg = TPG; p.a_row_bytes = 128; }
  CUtensorMap tmX, tmZ;
  {  // x, padded extent: {c, w, h, n}
    const uint64_t pix = (uint64_t)x->c * es, row = (uint64_t)t_wp(*x) * pix, img = (uint64_t)t_hp(*x) * row;
    uint64_t dims[4] = {(uint64_t)x->c, (uint64_t)t_wp(*x), (uint64_t)t_hp(*x), (uint64_t)x->n};
    uint64_t strides[3] = {pix, row, img};
    uint32_t box[4] = {(uint32_t)CBM, (uint32_t)p.pw, (uint32_t)p.ph, 1u};
    if (encode_tiled(&tmX, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, x->data, dims, strides, box,
                     row64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B))
      return -1;
  }
  {  // the window map over dz: {window element, q, y, n}, q = 0 is the window that starts k - 1 pixels left of the interior
    const uint64_t pix = (uint64_t)dz->c * es, row = (uint64_t)t_wp(*dz) * pix, img = (uint64_t)t_hp(*dz) * row;
    uint8_t* base = static_cast<uint8_t*>(dz->data) + (uint64_t)dz->halo * row + (uint64_t)(dz->halo - (k - 1)) * pix;
    uint64_t dims[4] = {(uint64_t)p.n_boxes * CB, (uint64_t)We, (uint64_t)dz->h, (uint64_t)dz->n};
    uint64_t strides[3] = {pix, row, img};
    uint32_t box[4] = {(uint32_t)CB, 8u, 8u, 1u};
    if (encode_tiled(&tmZ, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, base, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return -1;
  }
  const int smem_bytes = p.num_stages * p.stage_bytes + 1024;
  dim3 grid((unsigned)p.ksplit, (unsigned)slices, 1u);
  if (wgrad_patch_set_attr<1>()) return -1;
  launch_pdl(conv_wgrad_patch_kernel<1>, grid, 256, smem_bytes, static_cast<cudaStream_t>(stream), tmX, tmZ, p);
  UEGAN_CUDA(cudaGetLastError());
  return wg_reduce(dw_oihw, ws, p.dw_numel, p.ksplit, p.total_ktiles, cin_total, cin_first, cin, k,
                   static_cast<cudaStream_t>(stream));
}

extern "C" int uegan_wgrad_last_launches(void) { return uegan::g_wgrad_launches; }
