// HBM-bound staging kernels: coalesced vector loads / stores (position tables, dummy-token rows,
// pyramid level 0, layout conversion).  The big fp32 inputs are read by inproj.cu.
#include "kernels.cuh"
#include "ptx.cuh"

namespace fvtg {

// pos[b*Lv + i][c]: e = (i+1) / (len + 1e-6) * 2pi ; even c: sin(e / w_c), odd c: cos(e / w_c),
// w_c = 10000^(2*floor(c/2)/256).
__global__ void posenc_kernel(float* __restrict__ pos, const int* __restrict__ vlen, int B, int Lv) {
  const long long tt = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  // lanes run over consecutive rows of one 4-column group: 16-byte stores contiguous in the
  // tile-blocked layout
  const long long rows = static_cast<long long>(B) * Lv;
  const long long rows_pad = (rows + 127) & ~127ll;
  if (tt >= rows_pad * 64) return;
  const long long tile = tt / (64 * 128);
  const int rem = static_cast<int>(tt - tile * (64 * 128));
  const int c4 = rem >> 7, rin = rem & 127;
  const long long row = tile * 128 + rin;
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (row < rows) {
    const int b = static_cast<int>(row / Lv), i = static_cast<int>(row - static_cast<long long>(b) * Lv);
    const int len = vlen[b];
    if (i < len) {
      const float e = static_cast<float>(i + 1) / (static_cast<float>(len) + 1e-6f) * 6.283185307179586f;
      const float w0 = powf(10000.f, static_cast<float>(4 * c4) / 256.f);       // columns 4c4, 4c4+1
      const float w1 = powf(10000.f, static_cast<float>(4 * c4 + 2) / 256.f);   // columns 4c4+2, +3
      const float a0 = e / w0, a1 = e / w1;
      v = make_float4(sinf(a0), cosf(a0), sinf(a1), cosf(a1));
    }
  }
  *reinterpret_cast<float4*>(pos + (tile * 64 + c4) * 512 + rin * 4) = v;
}

// Compact table for a chunk whose videos all have the same true length vlen[0] (the pos rows
// depend on (position, length) only): pos_c[col/4][i][4], i < Lv; rows >= vlen[0] are zero.  Lanes
// of a warp read consecutive rows of one column group -> 16-byte loads stay contiguous.
__global__ void posenc_compact_kernel(float* __restrict__ pos, const int* __restrict__ vlen, int Lv) {
  const int tt = blockIdx.x * blockDim.x + threadIdx.x;
  if (tt >= Lv * 64) return;
  const int c4 = tt / Lv, i = tt - c4 * Lv;
  const int len = vlen[0];
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (i < len) {
    const float e = static_cast<float>(i + 1) / (static_cast<float>(len) + 1e-6f) * 6.283185307179586f;
    const float w0 = powf(10000.f, static_cast<float>(4 * c4) / 256.f);
    const float w1 = powf(10000.f, static_cast<float>(4 * c4 + 2) / 256.f);
    const float a0 = e / w0, a1 = e / w1;
    v = make_float4(sinf(a0), cosf(a0), sinf(a1), cosf(a1));
  }
  *reinterpret_cast<float4*>(pos + (static_cast<size_t>(c4) * Lv + i) * 4) = v;
}

int launch_posenc(cudaStream_t st, float* pos, const int* vlen, int B, int Lv, bool compact) {
  if (B * Lv <= 0) return FVTG_OK;
  if (compact) {
    ProfScope prof(st, PC_OTHER);
    posenc_compact_kernel<<<(Lv * 64 + 255) / 256, 256, 0, st>>>(pos, vlen, Lv);
    FVTG_LAUNCH_CHECK("posenc_compact_kernel");
    return FVTG_OK;
  }
  const long long rows_pad = ((static_cast<long long>(B) * Lv + 127) / 128) * 128;
  const long long threads = rows_pad * 64;
  ProfScope prof(st, PC_OTHER);
  posenc_kernel<<<static_cast<int>((threads + 255) / 256), 256, 0, st>>>(pos, vlen, B, Lv);
  FVTG_LAUNCH_CHECK("posenc_kernel");
  return FVTG_OK;
}

// thread = (video, dummy row, 4 columns)
__global__ void fill_dummy_kernel(const float* __restrict__ dtok, const float* __restrict__ dpos,
                                  float* __restrict__ X, bf16* __restrict__ Xb,
                                  bf16* __restrict__ XPb, float* __restrict__ pos_d, int B, int S,
                                  int nd) {
  const long long tt = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (tt < static_cast<long long>(S) * 64) {  // the [dummy_pos ‖ 0] table
    const int j = static_cast<int>(tt >> 6), c = static_cast<int>(tt & 63) * 4;
    float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
    if (j < nd) p = *reinterpret_cast<const float4*>(dpos + j * 256 + c);
    *reinterpret_cast<float4*>(pos_d + j * 256 + c) = p;
  }
  if (tt >= static_cast<long long>(B) * nd * 64) return;
  const int c = static_cast<int>(tt & 63) * 4;
  const long long rj = tt >> 6;
  const int b = static_cast<int>(rj / nd), j = static_cast<int>(rj - static_cast<long long>(b) * nd);
  const float4 tk = *reinterpret_cast<const float4*>(dtok + j * 256 + c);
  const float4 p = *reinterpret_cast<const float4*>(dpos + j * 256 + c);
  const size_t row = static_cast<size_t>(b) * S + j;
  *reinterpret_cast<float4*>(X + blk_off(row, c)) = tk;
  uint2 u, up;
  u.x = pack_bf16(tk.x, tk.y); u.y = pack_bf16(tk.z, tk.w);
  up.x = pack_bf16(tk.x + p.x, tk.y + p.y); up.y = pack_bf16(tk.z + p.z, tk.w + p.w);
  *reinterpret_cast<uint2*>(Xb + row * 256 + c) = u;
  *reinterpret_cast<uint2*>(XPb + row * 256 + c) = up;
}

int launch_fill_dummy(cudaStream_t st, const float* dtok, const float* dpos, float* X, bf16* Xb,
                      bf16* XPb, float* pos_d, int B, int S, int nd) {
  long long threads = static_cast<long long>(B) * nd * 64;
  if (threads < static_cast<long long>(S) * 64) threads = static_cast<long long>(S) * 64;
  ProfScope prof(st, PC_OTHER);
  fill_dummy_kernel<<<static_cast<int>((threads + 255) / 256), 256, 0, st>>>(dtok, dpos, X, Xb, XPb,
                                                                           pos_d, B, S, nd);
  FVTG_LAUNCH_CHECK("fill_dummy_kernel");
  return FVTG_OK;
}

// thread = (chain row, 8 channels)
__global__ void level0_kernel(const float* __restrict__ F, bf16* __restrict__ chain0,
                              bf16* __restrict__ H1, bf16* __restrict__ H2, int B, int Lv,
                              PyrGeo geo, int blocked) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int row = t >> 5, c8 = (t & 31) * 8;
  if (row >= B * geo.P0) return;
  const int b = row / geo.P0, i = row - b * geo.P0;
  const int vl = geo.vlen[b];
  uint4 o = make_uint4(0, 0, 0, 0);
  const bool valid = i < vl;
  if (valid) {
    const size_t frow = static_cast<size_t>(b) * Lv + i;
    float4 a, c;
    if (blocked) {
      a = __ldg(reinterpret_cast<const float4*>(F + blk_off(frow, c8)));
      c = __ldg(reinterpret_cast<const float4*>(F + blk_off(frow, c8 + 4)));
    } else {
      const float4* src = reinterpret_cast<const float4*>(F + frow * 256 + c8);
      a = __ldg(src);
      c = __ldg(src + 1);
    }
    o.x = pack_bf16(fmaxf(a.x, 0.f), fmaxf(a.y, 0.f));
    o.y = pack_bf16(fmaxf(a.z, 0.f), fmaxf(a.w, 0.f));
    o.z = pack_bf16(fmaxf(c.x, 0.f), fmaxf(c.y, 0.f));
    o.w = pack_bf16(fmaxf(c.z, 0.f), fmaxf(c.w, 0.f));
  }
  *reinterpret_cast<uint4*>(chain0 + static_cast<size_t>(row) * 256 + c8) = o;
  if (valid) {
    const size_t r1 = static_cast<size_t>(b) * geo.PH1 + geo.o1[0] + i;
    const size_t r2 = static_cast<size_t>(b) * geo.PH2 + geo.pad + i;
    *reinterpret_cast<uint4*>(H1 + r1 * 256 + c8) = o;
    *reinterpret_cast<uint4*>(H2 + r2 * 256 + c8) = o;
  }
  // The rows of H1 / H2 that NO producer ever writes - the pad rows around and between the levels, the
  // positions past this video's length in every level, the tail of the packed H2 row space - must read
  // as zero for the conv taps.  They are few (12 of 294 rows for a full-length 75-clip video), so the
  // threads of this video's first chain rows zero them here instead of two 77 MB memsets per forward.
  const uint4 z = make_uint4(0, 0, 0, 0);
  int nvalid = 0;
#pragma unroll
  for (int l = 0; l < FVTG_MAX_LEVELS; ++l) nvalid += l < geo.nlev ? vl >> l : 0;   // static indices: geo stays in param space
  const int nzero = (geo.PH1 - nvalid) + (geo.PH2 - nvalid);   // rows of this video nobody writes
  for (int k = i; k < nzero; k += geo.P0) {
    int q = k;
    bool done = false;
    // H1 segments: leading pad, then after every level's valid rows up to the next level's first row
    if (q < geo.pad) {
      *reinterpret_cast<uint4*>(H1 + (static_cast<size_t>(b) * geo.PH1 + q) * 256 + c8) = z;
      continue;
    }
    q -= geo.pad;
#pragma unroll
    for (int l = 0; l < FVTG_MAX_LEVELS; ++l) {
      if (l >= geo.nlev || done) continue;
      const int start = geo.o1[l] + (vl >> l);
      const int end = (l + 1 < FVTG_MAX_LEVELS && l + 1 < geo.nlev) ? geo.o1[l + 1 < FVTG_MAX_LEVELS ? l + 1 : l] : geo.PH1;
      if (q < end - start) {
        *reinterpret_cast<uint4*>(H1 + (static_cast<size_t>(b) * geo.PH1 + start + q) * 256 + c8) = z;
        done = true;
      } else {
        q -= end - start;
      }
    }
    if (done) continue;
    // H2 segments: leading pad, then everything behind the packed valid rows
    if (q < geo.pad) {
      *reinterpret_cast<uint4*>(H2 + (static_cast<size_t>(b) * geo.PH2 + q) * 256 + c8) = z;
      continue;
    }
    q -= geo.pad;
    const int tail0 = geo.pad + nvalid;
    if (q < geo.PH2 - tail0) {
      *reinterpret_cast<uint4*>(H2 + (static_cast<size_t>(b) * geo.PH2 + tail0 + q) * 256 + c8) = z;
      continue;
    }
    break;
  }
}

int launch_level0(cudaStream_t st, const float* F, bf16* chain0, bf16* H1, bf16* H2, int B, int Lv,
                  const PyrGeo& geo, bool blocked) {
  const long long threads = static_cast<long long>(B) * geo.P0 * 32;
  if (threads <= 0) return FVTG_OK;
  const int grid = static_cast<int>((threads + 255) / 256);
  ProfScope prof(st, PC_OTHER);
  level0_kernel<<<grid, 256, 0, st>>>(F, chain0, H1, H2, B, Lv, geo, blocked ? 1 : 0);
  FVTG_LAUNCH_CHECK("level0_kernel");
  return FVTG_OK;
}

// thread = (output row, 4 columns)
__global__ void unblock_kernel(const float* __restrict__ src, float* __restrict__ dst, int B,
                               int rows_in, int rows_out) {
  const long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long orow = t >> 6;
  const int c4 = static_cast<int>(t & 63) * 4;
  if (orow >= static_cast<long long>(B) * rows_out) return;
  const long long b = orow / rows_out, j = orow - b * rows_out;
  const float4 v = *reinterpret_cast<const float4*>(src + blk_off(static_cast<size_t>(b * rows_in + j), c4));
  *reinterpret_cast<float4*>(dst + orow * 256 + c4) = v;
}

int launch_unblock(cudaStream_t st, const float* src_blk, float* dst, int B, int rows_in, int rows_out) {
  const long long threads = static_cast<long long>(B) * rows_out * 64;
  if (threads <= 0) return FVTG_OK;
  ProfScope prof(st, PC_OTHER);
  unblock_kernel<<<static_cast<int>((threads + 255) / 256), 256, 0, st>>>(src_blk, dst, B, rows_in, rows_out);
  FVTG_LAUNCH_CHECK("unblock_kernel");
  return FVTG_OK;
}

}  // namespace fvtg
