"""Stage flags and tap ids of include/r2f_b200.h (kept importable without the CUDA library)."""
HALATION, MTF, GRAIN, GRAIN_BW, BURN = 0x01, 0x02, 0x04, 0x08, 0x10
SPATIAL = HALATION | MTF | GRAIN | BURN
TAPS = {"exposure": 1, "halation": 2, "density": 3, "mtf": 4, "grain": 5, "burn": 6, "rgb": 7}
