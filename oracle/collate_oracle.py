"""CPU restatement of the packed collation.  TEST INFRASTRUCTURE ONLY.

Follows /root/reference/src/datamodules/components/edm_dataset.py:187-216 (`_featurize_as_graph`: mask = charges > 0,
one_hot / charges as float32, coordinates of missing atoms zeroed) + PyG collation (concatenation, `batch` vector) and
/root/reference/src/datamodules/components/edm/utils.py:333-382 (`prepare_context`, global-property branch:
(p - mean) / mad repeated over the molecule's nodes, times the node mask), then keeps the rows with mask == True — the
packed batch.  Pinned against the reference's own functions by tests/golden/make_golden_collate.py.
"""
import numpy as np


def collate_packed(positions, charges, one_hot, idx):
    """padded positions [M,P,3], charges [M,P], one_hot [M,P,A]; idx [B] -> x [N,3], one_hot [N,A] f32, charges [N,1] f32,
    batch_index [N] i64, counts [B]."""
    xs, ohs, chs, bis, counts = [], [], [], [], []
    for k, m in enumerate(np.asarray(idx)):
        mask = np.asarray(charges[m]) > 0
        xs.append(np.asarray(positions[m], dtype=np.float32)[mask])
        ohs.append(np.asarray(one_hot[m]).astype(np.float32)[mask])
        chs.append(np.asarray(charges[m]).astype(np.float32)[mask][:, None])
        bis.append(np.full(int(mask.sum()), k, dtype=np.int64))
        counts.append(int(mask.sum()))
    return np.concatenate(xs), np.concatenate(ohs), np.concatenate(chs), np.concatenate(bis), np.asarray(counts)


def prepare_context(props, idx, batch_index, mean, mad):
    """props [C][M] per-molecule, mean/mad [C] -> context [N,C] = ((p - mean) / mad)[molecule of node], fp32 operations."""
    cols = []
    for c in range(len(props)):
        p = (np.asarray(props[c], dtype=np.float32)[np.asarray(idx)] - np.float32(mean[c])) / np.float32(mad[c])
        cols.append(p[np.asarray(batch_index)])
    return np.stack(cols, axis=1).astype(np.float32)
