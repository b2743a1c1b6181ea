"""CPU: the packed-collation oracle (oracle/collate_oracle.py) against what the unmodified reference makes of the same batch
(tests/golden/collate.pt from tests/golden/make_golden_collate.py): `_featurize_as_graph` + collation + `prepare_context`,
restricted to mask == True rows — bit-exact."""
import os

import numpy as np
import torch

import collate_oracle as CO
from conftest import GOLDEN


def test_packed_batch_equals_reference_batch_restricted_to_mask():
    fx = torch.load(os.path.join(GOLDEN, "collate.pt"), weights_only=False)
    x, oh, ch, bi, counts = CO.collate_packed(fx["positions"].numpy(), fx["charges"].numpy(), fx["one_hot"].numpy(),
                                              fx["idx"].numpy())
    ref = fx["ref"]
    assert np.array_equal(x, ref["x"].numpy()) and np.array_equal(oh, ref["one_hot"].numpy())
    assert np.array_equal(ch, ref["charges"].numpy()) and np.array_equal(bi, ref["batch"].numpy())
    assert counts.sum() == ref["present_rows"] < ref["padded_rows"]
    ctx = CO.prepare_context([fx["alpha"].numpy(), fx["mu"].numpy()], fx["idx"].numpy(), bi,
                             [fx["norms"]["alpha"]["mean"].item(), fx["norms"]["mu"]["mean"].item()],
                             [fx["norms"]["alpha"]["mad"].item(), fx["norms"]["mu"]["mad"].item()])
    assert np.array_equal(ctx, ref["context"].numpy())
