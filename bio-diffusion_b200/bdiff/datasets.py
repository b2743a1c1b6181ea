"""Dataset statistics the sampler needs: the number-of-atoms histograms of QM9 and GEOM-Drugs (with hydrogens).

These are the `n_nodes` tables of the reference's dataset descriptions (data, not code:
src/datamodules/components/edm/datasets_config.py:38-40 QM9_WITH_H, :116-141 GEOM_WITH_H) and the categorical
sampler over them (`NumNodesDistribution.sample`, src/models/__init__.py:264-297).
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

QM9_N_NODES: Dict[int, int] = {
    22: 3393, 17: 13025, 23: 4848, 21: 9970, 19: 13832, 20: 9482, 16: 10644, 13: 3060, 15: 7796, 25: 1506, 18: 13364,
    12: 1689, 11: 807, 24: 539, 14: 5136, 26: 48, 7: 16, 10: 362, 8: 49, 9: 124, 27: 266, 4: 4, 29: 25, 6: 9, 5: 5, 3: 1}

GEOM_N_NODES: Dict[int, int] = {
    3: 1, 4: 3, 5: 9, 6: 2, 7: 8, 8: 23, 9: 23, 10: 50, 11: 109, 12: 168, 13: 280, 14: 402, 15: 583, 16: 597,
    17: 949, 18: 1284, 19: 1862, 20: 2674, 21: 3599, 22: 6109, 23: 8693, 24: 13604, 25: 17419, 26: 25672,
    27: 31647, 28: 43809, 29: 56697, 30: 70400, 31: 82655, 32: 104100, 33: 122776, 34: 140834, 35: 164888,
    36: 185451, 37: 194541, 38: 218549, 39: 231232, 40: 243300, 41: 253349, 42: 268341, 43: 272081,
    44: 276917, 45: 276839, 46: 274747, 47: 272126, 48: 262709, 49: 250157, 50: 244781, 51: 228898,
    52: 215338, 53: 203728, 54: 191697, 55: 180518, 56: 163843, 57: 152055, 58: 136536, 59: 120393,
    60: 107292, 61: 94635, 62: 83179, 63: 68384, 64: 61517, 65: 48867, 66: 37685, 67: 32859, 68: 27367,
    69: 20981, 70: 18699, 71: 14791, 72: 11921, 73: 9933, 74: 9037, 75: 6538, 76: 6374, 77: 4036, 78: 4189,
    79: 3842, 80: 3277, 81: 2925, 82: 1843, 83: 2060, 84: 1394, 85: 1514, 86: 1357, 87: 1346, 88: 999,
    89: 300, 90: 390, 91: 510, 92: 510, 93: 240, 94: 721, 95: 360, 96: 360, 97: 390, 98: 330, 99: 540,
    100: 258, 101: 210, 102: 60, 103: 180, 104: 206, 105: 60, 106: 390, 107: 180, 108: 180, 109: 150,
    110: 120, 111: 360, 112: 120, 113: 210, 114: 60, 115: 30, 116: 210, 117: 270, 118: 450, 119: 240,
    120: 228, 121: 120, 122: 30, 123: 420, 124: 240, 125: 210, 126: 158, 127: 180, 128: 60, 129: 30,
    130: 120, 131: 30, 132: 120, 133: 60, 134: 240, 135: 169, 136: 240, 137: 30, 138: 270, 139: 180,
    140: 270, 141: 150, 142: 60, 143: 60, 144: 240, 145: 180, 146: 150, 147: 150, 148: 90, 149: 90,
    151: 30, 152: 60, 155: 90, 159: 30, 160: 60, 165: 30, 171: 30, 175: 30, 176: 60, 181: 30}


def sample_num_nodes(histogram: Dict[int, int], n_samples: int, seed: Optional[int] = None) -> torch.Tensor:
    """NumNodesDistribution.sample: categorical over the histogram (same torch.multinomial draw as the reference's
    torch.distributions.Categorical; pass `seed` for a private generator, else the global CPU generator is used)."""
    keys = torch.tensor(list(histogram.keys()))
    prob = torch.tensor([float(histogram[int(k)]) for k in keys])
    prob = prob / prob.sum()
    g = torch.Generator().manual_seed(seed) if seed is not None else None
    idx = torch.multinomial(prob, n_samples, replacement=True, generator=g)
    return keys[idx]
