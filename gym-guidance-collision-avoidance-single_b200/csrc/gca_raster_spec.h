// gca_raster_spec.h - the per-sample arithmetic of the image observation, shared by the CUDA
// rasteriser and the CPU oracle so that both evaluate literally the same f32 operations
// (each one a single round-to-nearest multiply / add, see gca_math.h).  DESIGN.md section 4.5.
#ifndef GCA_RASTER_SPEC_H_
#define GCA_RASTER_SPEC_H_

#include "gca_math.h"

#define GCA_SPRITE 32
#define GCA_SPRITE_HALF 16.0f
#define GCA_SPRITE_REACH 23.0f   /* > 16 * sqrt(2): conservative half-extent of a rotated sprite */

typedef struct gca_sprite_pose {
  float cx, cy;   /* centre, world units (origin bottom-left) */
  float rc, rs;   /* cos / sin of the quad rotation (heading - pi/2) */
  int tex;        /* 0 ownship, 1 goal, 2 intruder */
} gca_sprite_pose;

GCA_HD int gca_clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// Blend sprite `sp` over the 8-bit colour (r, g, b) of the sample whose centre is (wx, wy).
// tex: uint8 [32][32][4], rows top->bottom.  Returns 1 if the sample is inside the quad.
GCA_HD int gca_raster_sample(const gca_sprite_pose sp, const unsigned char* tex, float wx, float wy, int* r, int* g,
                             int* b) {
  const float dx = GCA_FSUBF(wx, sp.cx), dy = GCA_FSUBF(wy, sp.cy);
  // local = R(-theta) (w - pos)
  const float lx = GCA_FADDF(GCA_FMULF(sp.rc, dx), GCA_FMULF(sp.rs, dy));
  const float ly = GCA_FSUBF(GCA_FMULF(sp.rc, dy), GCA_FMULF(sp.rs, dx));
  if (!(lx >= -GCA_SPRITE_HALF && lx < GCA_SPRITE_HALF && ly >= -GCA_SPRITE_HALF && ly < GCA_SPRITE_HALF)) return 0;
  // GL_LINEAR with clamp-to-edge: texel centres at integer + 0.5
  const float tu = GCA_FADDF(lx, 15.5f), tv = GCA_FADDF(ly, 15.5f);
  const float fu0 = (float)(int)tu - (tu < (float)(int)tu ? 1.0f : 0.0f);    /* floorf without libm */
  const float fv0 = (float)(int)tv - (tv < (float)(int)tv ? 1.0f : 0.0f);
  const int iu = (int)fu0, iv = (int)fv0;
  const float fu = GCA_FSUBF(tu, fu0), fv = GCA_FSUBF(tv, fv0);
  const int u0 = gca_clampi(iu, 0, 31), u1 = gca_clampi(iu + 1, 0, 31);
  const int v0 = gca_clampi(iv, 0, 31), v1 = gca_clampi(iv + 1, 0, 31);
  const unsigned char* t00 = tex + ((31 - v0) * GCA_SPRITE + u0) * 4;        /* image row 0 is the top */
  const unsigned char* t10 = tex + ((31 - v0) * GCA_SPRITE + u1) * 4;
  const unsigned char* t01 = tex + ((31 - v1) * GCA_SPRITE + u0) * 4;
  const unsigned char* t11 = tex + ((31 - v1) * GCA_SPRITE + u1) * 4;
  const float gu = GCA_FSUBF(1.0f, fu), gv = GCA_FSUBF(1.0f, fv);
#define GCA_BILERP(k)                                                                                    \
  GCA_FADDF(GCA_FMULF(GCA_FADDF(GCA_FMULF((float)t00[k], gu), GCA_FMULF((float)t10[k], fu)), gv),        \
            GCA_FMULF(GCA_FADDF(GCA_FMULF((float)t01[k], gu), GCA_FMULF((float)t11[k], fu)), fv))
  // GL_SRC_ALPHA, GL_ONE_MINUS_SRC_ALPHA into an 8-bit framebuffer (quantised after every draw)
  const float a255 = GCA_BILERP(3);
  if (a255 == 0.0f) return 1;                         /* fully transparent: the blend is the identity */
  const float alpha = GCA_FMULF(a255, 0.003921568859368563f);              /* RN(1/255) */
  const float beta = GCA_FSUBF(1.0f, alpha);
  {
    const float v = GCA_FADDF(GCA_FMULF(GCA_BILERP(0), alpha), GCA_FMULF((float)*r, beta));
    *r = gca_clampi((int)GCA_RINTF(v), 0, 255);
  }
  {
    const float v = GCA_FADDF(GCA_FMULF(GCA_BILERP(1), alpha), GCA_FMULF((float)*g, beta));
    *g = gca_clampi((int)GCA_RINTF(v), 0, 255);
  }
  {
    const float v = GCA_FADDF(GCA_FMULF(GCA_BILERP(2), alpha), GCA_FMULF((float)*b, beta));
    *b = gca_clampi((int)GCA_RINTF(v), 0, 255);
  }
#undef GCA_BILERP
  return 1;
}

// cv2.cvtColor(COLOR_RGB2GRAY) on uint8: fixed point, SURVEY a12
GCA_HD int gca_gray_u8(int r, int g, int b) { return (9798 * r + 19235 * g + 3735 * b + 16384) >> 15; }

// cv2.resize(INTER_AREA) by an integer factor 4: mean of 16 samples, round half to even
GCA_HD int gca_area16_u8(int sum) { return (sum + 7 + ((sum >> 4) & 1)) >> 4; }

#endif  // GCA_RASTER_SPEC_H_
