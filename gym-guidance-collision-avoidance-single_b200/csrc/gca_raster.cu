// gca_raster.cu - image observation of SingleAircraftStackEnv on the device (sm_100a).
//
// Reference: PKG/SingleAircraftStackEnv.py:179-214 (render: 800x800 RGB GL frame) and :104-108
// (preprocess_frame: RGB2GRAY + INTER_AREA / 4).  One CTA per environment.  The 1.9 MB full
// resolution frame never exists: >= 95 % of the 200x200 output is background, so
//   1. the env's sprite poses (ownship, goal, intruders, in draw order) go to shared memory,
//   2. every sprite ORs its bit into the <= 3x3 cells (8x8 output pixels each) its reach touches,
//   3. the plane is filled with white (16-byte stores),
//   4. one warp per sprite walks the <= 13x13 output pixels of that sprite's bounding box; the 16
//      samples of a pixel are blended against the cell's sprites in bit (= draw) order with the
//      shared per-sample arithmetic of gca_raster_spec.h (fully transparent texels exit early),
//      then gray -> 4x4 area mean (round half to even); only touched pixels are rewritten.
// Output bytes: 40 000 per env-step.
#include "gca_launch.h"
#include "gca_raster_spec.h"

namespace gca {

constexpr int kRasterThreads = 256;
constexpr int kMaxSprites = GCA_RASTER_MAX_INTRUDERS + 2;   // 128 -> 4 mask words per cell

struct RasterArgs {
  DevState s;
  int faithful;
  int W, H;                 // full-resolution canvas
  int ow, oh;               // output size (W/4, H/4)
  int cells_x, cells_y;     // 8x8-output-pixel cells
  const uint8_t* sprites;
  uint8_t* frames;
  long long env_stride, plane_stride;
  int n_planes, slot;
  const uint8_t* clear_mask;
};

__global__ void __launch_bounds__(kRasterThreads) raster_kernel(const RasterArgs a) {
  extern __shared__ __align__(16) uint8_t rsm[];
  const DevState& s = a.s;
  const int env = blockIdx.x;
  const int tid = threadIdx.x;
  const int n_sprites = 2 + s.N;
  const int n_cells = a.cells_x * a.cells_y;
  uint8_t* tex = rsm;                                                  // [3][32][32][4]
  gca_sprite_pose* pose = reinterpret_cast<gca_sprite_pose*>(tex + 3 * 32 * 32 * 4);
  uint4* cell_mask = reinterpret_cast<uint4*>(pose + kMaxSprites);     // [n_cells]

  for (int i = tid; i < 3 * 32 * 32; i += kRasterThreads)
    reinterpret_cast<uint32_t*>(tex)[i] = reinterpret_cast<const uint32_t*>(a.sprites)[i];
  for (int i = tid; i < n_cells; i += kRasterThreads) cell_mask[i] = make_uint4(0u, 0u, 0u, 0u);

  // ---- 1. sprite poses in draw order: ownship, goal, intruders (PKG/SingleAircraftStackEnv.py:192-212)
  if (tid < n_sprites) {
    gca_sprite_pose p;
    if (tid == 0) {
      const float2 pos = s.own_pos[env];
      const double2 hs = s.own_hs[env];
      double sn, cs;
      gca_sincos(hs.x, &sn, &cs);
      p.cx = pos.x; p.cy = pos.y;
      p.rc = (float)sn;                                                // cos(h - pi/2) = sin h
      p.rs = -(float)cs;                                               // sin(h - pi/2) = -cos h
      p.tex = 0;
    } else if (tid == 1) {
      const double2 g = s.goal[env];
      p.cx = (float)g.x; p.cy = (float)g.y; p.rc = 1.0f; p.rs = 0.0f; p.tex = 1;
    } else {
      const int i = tid - 2;
      float px, py;
      const uint8_t* plane = s.ipos + (size_t)(s.counters[env].z & 1) * s.pos_plane;   // the env's current positions
      if (a.faithful) {
        const double2 q = *reinterpret_cast<const double2*>(plane + ipos_offset(s, true, (size_t)env, i));
        px = (float)q.x; py = (float)q.y;
      } else {
        const float2 q = *reinterpret_cast<const float2*>(plane + ipos_offset(s, false, (size_t)env, i));
        px = q.x; py = q.y;
      }
      const float2 v = *reinterpret_cast<const float2*>(s.ivel + ivel_offset(s, (size_t)env, i));
      // the intruder's heading is constant for life (:207-211); its direction is that of the velocity
      const float len = __fsqrt_rn(__fadd_rn(__fmul_rn(v.x, v.x), __fmul_rn(v.y, v.y)));
      const float ch = __fdiv_rn(v.x, len), sh = __fdiv_rn(v.y, len);
      p.cx = px; p.cy = py; p.rc = sh; p.rs = -ch; p.tex = 2;
    }
    pose[tid] = p;
  }
  // VecFrameStack: a finished env starts from an all-zero stack (vec_frame_stack.py:19-23)
  uint8_t* env_base = a.frames + (long long)env * a.env_stride;
  if (a.clear_mask && a.clear_mask[env]) {
    const int words = a.ow * a.oh / 4;
    for (int pl = 0; pl < a.n_planes; ++pl) {
      if (pl == a.slot) continue;
      uint32_t* dst = reinterpret_cast<uint32_t*>(env_base + (long long)pl * a.plane_stride);
      for (int i = tid; i < words; i += kRasterThreads) dst[i] = 0u;
    }
  }
  __syncthreads();

  // ---- 2. bin the sprites: cell (cx, cy) = 32x32 full-resolution pixels, y counted from the top row
  if (tid < n_sprites) {
    const gca_sprite_pose p = pose[tid];
    const float ytop = (float)a.H - p.cy;
    const int x0 = (int)floorf((p.cx - GCA_SPRITE_REACH) * (1.0f / 32.0f));
    const int x1 = (int)floorf((p.cx + GCA_SPRITE_REACH) * (1.0f / 32.0f));
    const int y0 = (int)floorf((ytop - GCA_SPRITE_REACH) * (1.0f / 32.0f));
    const int y1 = (int)floorf((ytop + GCA_SPRITE_REACH) * (1.0f / 32.0f));
    for (int cy = max(y0, 0); cy <= min(y1, a.cells_y - 1); ++cy)
      for (int cx = max(x0, 0); cx <= min(x1, a.cells_x - 1); ++cx)
        atomicOr(reinterpret_cast<unsigned int*>(&cell_mask[cy * a.cells_x + cx]) + (tid >> 5), 1u << (tid & 31));
  }
  __syncthreads();

  // ---- 3. background: the whole plane is white; ---- 4. sprite-centric pass over covered pixels only
  uint8_t* out = env_base + (long long)a.slot * a.plane_stride;
  {
    uint4* o4 = reinterpret_cast<uint4*>(out);
    const int n16 = a.ow * a.oh / 16;
    const uint4 white = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
    for (int i = tid; i < n16; i += kRasterThreads) o4[i] = white;
  }
  __syncthreads();
  // One warp per sprite: lanes walk the output pixels of the sprite's bounding box.  A pixel's value
  // depends on ALL sprites of its cell (blended in draw order), so pixels shared by two boxes are
  // computed twice and written twice with the same byte - benign.
  const int lane = tid & 31, warp = tid >> 5;
  for (int k0 = warp; k0 < n_sprites; k0 += kRasterThreads / 32) {
    const gca_sprite_pose me = pose[k0];
    const float ytop = (float)a.H - me.cy;
    const int bx0 = max((int)floorf((me.cx - GCA_SPRITE_REACH) * 0.25f), 0);
    const int bx1 = min((int)floorf((me.cx + GCA_SPRITE_REACH) * 0.25f), a.ow - 1);
    const int by0 = max((int)floorf((ytop - GCA_SPRITE_REACH) * 0.25f), 0);
    const int by1 = min((int)floorf((ytop + GCA_SPRITE_REACH) * 0.25f), a.oh - 1);
    const int nx = bx1 - bx0 + 1, ny = by1 - by0 + 1;
    if (nx <= 0 || ny <= 0) continue;
    for (int p = lane; p < nx * ny; p += 32) {
      const int oy = by0 + p / nx, ox = bx0 + p % nx;
      const uint4 m = cell_mask[(oy >> 3) * a.cells_x + (ox >> 3)];
      const uint32_t words[4] = {m.x, m.y, m.z, m.w};
      // sprites of the cell that can reach this pixel at all (centre distance test, conservative)
      const float pcx = (float)(4 * ox) + 2.0f, pcy = (float)a.H - ((float)(4 * oy) + 2.0f);
      uint32_t live[4];
      bool any = false;
      for (int w = 0; w < 4; ++w) {
        uint32_t bits = words[w], keep = 0u;
        while (bits) {
          const int j = __ffs(bits) - 1;
          bits &= bits - 1;
          const gca_sprite_pose sp = pose[w * 32 + j];
          if (fabsf(pcx - sp.cx) <= GCA_SPRITE_REACH + 2.0f && fabsf(pcy - sp.cy) <= GCA_SPRITE_REACH + 2.0f) keep |= 1u << j;
        }
        live[w] = keep;
        any |= keep != 0u;
      }
      if (!any) continue;
      int sum = 0;
      bool touched = false;
      for (int sy = 0; sy < 4; ++sy) {
        const float wy = (float)a.H - ((float)(4 * oy + sy) + 0.5f);
        for (int sx = 0; sx < 4; ++sx) {
          const float wx = (float)(4 * ox + sx) + 0.5f;
          int r = 255, g = 255, b = 255;                                  // white clear
          for (int w = 0; w < 4; ++w) {
            uint32_t bits = live[w];
            while (bits) {
              const int k = w * 32 + __ffs(bits) - 1;
              bits &= bits - 1;
              const gca_sprite_pose sp = pose[k];
              touched |= gca_raster_sample(sp, tex + sp.tex * (32 * 32 * 4), wx, wy, &r, &g, &b) != 0;
            }
          }
          sum += gca_gray_u8(r, g, b);
        }
      }
      if (touched) out[oy * a.ow + ox] = (uint8_t)gca_area16_u8(sum);
    }
  }
}

cudaError_t launch_raster(const DevState& s, bool faithful, int W, int H, const uint8_t* sprites, uint8_t* frames,
                          long long env_stride, long long plane_stride, int n_planes, int slot,
                          const uint8_t* clear_mask, cudaStream_t st) {
  RasterArgs a{};
  a.s = s; a.faithful = faithful ? 1 : 0; a.W = W; a.H = H; a.ow = W / 4; a.oh = H / 4;
  a.cells_x = (a.ow + 7) / 8; a.cells_y = (a.oh + 7) / 8;
  a.sprites = sprites; a.frames = frames; a.env_stride = env_stride; a.plane_stride = plane_stride;
  a.n_planes = n_planes; a.slot = slot; a.clear_mask = clear_mask;
  const size_t smem = 3 * 32 * 32 * 4 + sizeof(gca_sprite_pose) * kMaxSprites + sizeof(uint4) * (size_t)a.cells_x * a.cells_y;
  cudaError_t e = cudaFuncSetAttribute(raster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  raster_kernel<<<(unsigned)s.B, kRasterThreads, smem, st>>>(a);
  return cudaGetLastError();
}

}  // namespace gca
