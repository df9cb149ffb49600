/* gca_oracle_raster.c - CPU restatement of the StackEnv image observation
 * (PKG/SingleAircraftStackEnv.py:179-214 render(), :104-108 preprocess_frame()).
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT.  PARITY UNPINNED for the GL half: the reference's render()
 * needs pyglet + an OpenGL display and cannot run in the build container, and no reference test
 * pins a pixel; this restatement, written from the reference's render() code and
 * gym's rendering semantics, IS the specification.  The cv2 half (gray, area
 * resize) is pinned against the real cv2 in tests/test_raster_cpu.py.
 *
 * Deliberately the naive algorithm: a full 800x800x3 framebuffer, painter's order, every pixel
 * of every sprite's bounding box, then gray + 4x4 mean - only the per-sample arithmetic
 * (gca_raster_spec.h) is shared with the CUDA kernel, whose cell binning / gather structure
 * it therefore cross-checks.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "gca_oracle.h"
#include "gca_raster_spec.h"

static void draw(unsigned char* fb, int W, int H, const gca_sprite_pose sp, const unsigned char* tex) {
  const float ytop = (float)H - sp.cy;
  int x0 = (int)floorf(sp.cx - GCA_SPRITE_REACH), x1 = (int)ceilf(sp.cx + GCA_SPRITE_REACH);
  int y0 = (int)floorf(ytop - GCA_SPRITE_REACH), y1 = (int)ceilf(ytop + GCA_SPRITE_REACH);
  if (x0 < 0) x0 = 0;
  if (y0 < 0) y0 = 0;
  if (x1 > W - 1) x1 = W - 1;
  if (y1 > H - 1) y1 = H - 1;
  for (int Y = y0; Y <= y1; ++Y)
    for (int X = x0; X <= x1; ++X) {
      unsigned char* px = fb + ((size_t)Y * W + X) * 3;
      int r = px[0], g = px[1], b = px[2];
      const float wx = (float)X + 0.5f, wy = (float)H - ((float)Y + 0.5f);
      if (gca_raster_sample(sp, tex, wx, wy, &r, &g, &b)) {
        px[0] = (unsigned char)r; px[1] = (unsigned char)g; px[2] = (unsigned char)b;
      }
    }
}

/* frames: uint8 [B][H/4][W/4]; rgb (nullable): uint8 [B][H][W][3] full-resolution frames */
int gca_oracle_raster(const gca_config* cfg, const gca_oracle_batch* b, const unsigned char* sprites,
                      unsigned char* frames, unsigned char* rgb) {
  const int W = (int)cfg->window_width, H = (int)cfg->window_height, n = b->n_intr;
  const int ow = W / 4, oh = H / 4;
  unsigned char* fb = (unsigned char*)malloc((size_t)W * H * 3);
  if (!fb) return GCA_ERR_ALLOC;
  for (int e = 0; e < b->n_envs; ++e) {
    memset(fb, 255, (size_t)W * H * 3);                                   /* white clear */
    gca_sprite_pose p;
    double s, c;
    gca_oracle_sincos(b->st.own_hs[2 * e], b->trig, &s, &c);              /* ownship: rotation heading - pi/2 */
    p.cx = b->st.own_pos[2 * e]; p.cy = b->st.own_pos[2 * e + 1];
    p.rc = (float)s; p.rs = -(float)c; p.tex = 0;
    draw(fb, W, H, p, sprites);
    p.cx = (float)b->st.goal[2 * e]; p.cy = (float)b->st.goal[2 * e + 1]; /* goal: rotation 0 */
    p.rc = 1.0f; p.rs = 0.0f; p.tex = 1;
    draw(fb, W, H, p, sprites + 32 * 32 * 4);
    for (int i = 0; i < n; ++i) {                                         /* intruders in list order */
      const size_t k = (size_t)e * n + i;
      const float vx = b->st.ivel[2 * k], vy = b->st.ivel[2 * k + 1];
      const float len = sqrtf(vx * vx + vy * vy);
      p.cx = (float)b->st.ipos[2 * k]; p.cy = (float)b->st.ipos[2 * k + 1];
      p.rc = len > 0.0f ? vy / len : 0.0f; p.rs = len > 0.0f ? -(vx / len) : -1.0f; p.tex = 2;   /* (no direction: heading 0) */
      draw(fb, W, H, p, sprites + 2 * 32 * 32 * 4);
    }
    if (rgb) memcpy(rgb + (size_t)e * W * H * 3, fb, (size_t)W * H * 3);
    unsigned char* out = frames + (size_t)e * ow * oh;
    for (int oy = 0; oy < oh; ++oy)
      for (int ox = 0; ox < ow; ++ox) {
        int sum = 0;
        for (int sy = 0; sy < 4; ++sy)
          for (int sx = 0; sx < 4; ++sx) {
            const unsigned char* px = fb + ((size_t)(4 * oy + sy) * W + 4 * ox + sx) * 3;
            sum += gca_gray_u8(px[0], px[1], px[2]);
          }
        out[oy * ow + ox] = (unsigned char)gca_area16_u8(sum);
      }
  }
  free(fb);
  return GCA_OK;
}
