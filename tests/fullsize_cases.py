"""The full-size parity cases shared by tests/golden/make_golden_fullsize.py (cv2 -> fixtures) and the tests that check the
oracle and the CUDA path against those fixtures.  Pure NumPy: importable without cv2."""
from zenslam_b200 import synthetic as syn

CASES = {
    #  name    (w, h)        seed  cell      thr subpix  LK settings (win, max_level)            max tracks per job, extra (random / border) tracks
    "c2":    ((752, 480),   2001, (16, 16), 10, False, [((31, 31), 3)],                  0,    300),
    "tumvi": ((752, 480),   2002, (64, 64), 1,  True,  [((63, 63), 4)],                  0,    1300),
    "c4":    ((1280, 1024), 2004, (16, 16), 10, False, [((31, 31), 3), ((63, 63), 4)],   2500, 200),
    "c5":    ((3840, 2160), 2005, (32, 32), 10, False, [((31, 31), 3)],                  4000, 200),
}


def frames(name):
    """-> left_0, right_0, left_1 of the case (u8): a stereo pair and the next left frame after a sub-pixel move"""
    (w, h), seed = CASES[name][0], CASES[name][1]
    base = syn.base_texture(w, h, seed)
    return (syn.crop(base, w, h, 0, 0), syn.crop(base, w, h, 6, 0), syn.crop(base, w, h, 4.25, -3.5))
