"""Sprite textures of the image observation (SingleAircraftStackEnv).

The reference draws three 32x32 RGBA sprites (PKG/images/{aircraft,goal,intruder}.png,
PKG/SingleAircraftStackEnv.py:193-212).  Those files are the reference's assets and are not
shipped with the product: `resolve_sprites()` loads the ORIGINAL PNGs when it is told where a checkout of
the reference (or just its images directory) is - constructor argument `sprite_dir=` or the environment
variables GCA_SPRITE_DIR / GCA_REFERENCE - and only then is the picture the reference's picture
(pinned to an independent renderer in tests/test_gpu_raster.py).  Without one it falls back, with a
warning, to `default_sprites()`: look-alikes built procedurally (an aircraft silhouette nose-up in
yellow / red, a five-pointed green star, anti-aliased alpha) - same geometry, different texels.
Layout: uint8 [3, 32, 32, 4] = (ownship, goal, intruder) x rows top->bottom x columns x RGBA.
"""
import os
import warnings

import numpy as np

SIZE = 32
OWNSHIP, GOAL, INTRUDER = 0, 1, 2


def _coverage(inside, ss=8):
    """Fraction of each texel covered by the shape `inside(x, y)` (x right, y up, texel units)."""
    k = (np.arange(SIZE * ss) + 0.5) / ss
    x, y = np.meshgrid(k, SIZE - k)                      # row 0 is the top of the image
    m = inside(x, y).astype(np.float64)
    return m.reshape(SIZE, ss, SIZE, ss).mean((1, 3))


def _aircraft(x, y):
    cx = 16.0
    body = (np.abs(x - cx) <= 2.2) & (y >= 1.0) & (y <= 31.5)
    nose = (np.abs(x - cx) <= 2.2 * (31.9 - y) / 2.0) & (y > 29.5)
    # swept main wing: leading edge falls from the fuselage (y = 19) to the tips (y = 11), chord shrinks outward
    t = np.abs(x - cx) / 16.0
    wing = (t <= 1.0) & (y <= 19.0 - 8.0 * t) & (y >= 12.5 - 3.5 * t) & (y >= 8.0)
    tail = (np.abs(x - cx) <= 6.0) & (y <= 4.5 - 0.25 * np.abs(x - cx)) & (y >= 1.0)
    return (body & ~(y > 29.5)) | nose | wing | tail


def _star(x, y):
    cx, cy, R, r = 16.0, 15.0, 16.0, 6.4
    ang = np.arctan2(y - cy, x - cx) - np.pi / 2
    rad = np.hypot(x - cx, y - cy)
    a = np.mod(ang, 2 * np.pi / 5)
    a = np.minimum(a, 2 * np.pi / 5 - a)                 # angle to the nearest outer tip
    # boundary of a 5-pointed star in polar form between a tip (radius R) and a notch (radius r)
    half = np.pi / 5
    edge = R * r * np.sin(half) / (R * np.sin(a) + r * np.sin(half - a) + 1e-12)
    return rad <= edge


def default_sprites():
    out = np.zeros((3, SIZE, SIZE, 4), np.uint8)
    for idx, shape, rgb in ((OWNSHIP, _aircraft, (230, 219, 0)), (GOAL, _star, (73, 192, 107)),
                            (INTRUDER, _aircraft, (215, 26, 33))):
        cov = _coverage(shape)
        out[idx, :, :, :3] = np.array(rgb, np.uint8)
        out[idx, :, :, 3] = np.rint(cov * 255).astype(np.uint8)
    return out


def load_sprites(image_dir):
    """Load aircraft.png / goal.png / intruder.png of a reference checkout (needs Pillow)."""
    from PIL import Image
    out = np.zeros((3, SIZE, SIZE, 4), np.uint8)
    for idx, name in ((OWNSHIP, "aircraft.png"), (GOAL, "goal.png"), (INTRUDER, "intruder.png")):
        im = Image.open(os.path.join(image_dir, name)).convert("RGBA")
        a = np.asarray(im, np.uint8)
        if a.shape != (SIZE, SIZE, 4):
            raise ValueError("%s must be %dx%d" % (name, SIZE, SIZE))
        out[idx] = a
    return out


_IMAGES_SUBDIR = os.path.join("gym_guidance_collision_avoidance_single", "envs", "images")


def find_sprite_dir(sprite_dir=None):
    """Directory holding the reference's aircraft.png / goal.png / intruder.png, or None.  `sprite_dir` (argument),
    then $GCA_SPRITE_DIR, then $GCA_REFERENCE; each may be the images directory itself or the root of a checkout."""
    for cand in (sprite_dir, os.environ.get("GCA_SPRITE_DIR"), os.environ.get("GCA_REFERENCE")):
        if not cand:
            continue
        for d in (cand, os.path.join(cand, _IMAGES_SUBDIR)):
            if os.path.isfile(os.path.join(d, "aircraft.png")):
                return d
        if cand is sprite_dir:
            raise FileNotFoundError("sprite_dir=%r holds no aircraft.png (nor %s under it)" % (cand, _IMAGES_SUBDIR))
    return None


def resolve_sprites(sprites=None, sprite_dir=None):
    """The texture array the rasteriser gets: an explicit array, else the reference's PNGs (find_sprite_dir), else the
    look-alikes - with a warning, because the frames then differ from the reference's by construction."""
    if sprites is not None:
        sp = np.ascontiguousarray(sprites, np.uint8)
        if sp.shape != (3, SIZE, SIZE, 4):
            raise ValueError("sprites must be uint8 [3, 32, 32, 4]")
        return sp
    d = find_sprite_dir(sprite_dir)
    if d is not None:
        return load_sprites(d)
    warnings.warn("SingleAircraftStackEnv: the reference's sprite PNGs were not found (pass sprite_dir= or set "
                  "GCA_SPRITE_DIR / GCA_REFERENCE to a checkout of the reference); drawing procedural look-alikes - "
                  "geometry, draw order and post-processing are the reference's, the texels are not", RuntimeWarning,
                  stacklevel=3)
    return default_sprites()
