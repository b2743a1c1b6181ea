"""Hand-derived backward pass of the GCPNet denoiser (no autograd) — TEST INFRASTRUCTURE ONLY.

What a CUDA backward of the hot path has to compute, written out operation by operation so that every formula a kernel
will implement is pinned before the kernel exists (BASELINE config 5).  `denoiser_backward` returns
the gradient of  sum(net_out * d_net_out)  with respect to every parameter of `gcpnet_oracle.denoiser_forward`
(same state-dict keys); tests/test_oracle_backward.py checks it against torch.autograd through the forward oracle, and
tests/test_oracle_golden.py::test_training_gradients_match_reference pins that autograd path against loss.backward() of
the unmodified reference.  Forward formulas: gcpnet_oracle.py (each citing /root/reference/src/models/components/
gcpnet.py); the derivatives below follow them line by line.

Structure of the reverse sweep (the forward kernels' order, reversed):
    net_out -> [centralize, projection GCP] -> L x interaction layer {position GCP, residual/mask, feed-forward GCP,
    message passing {row-sum, attention gate, 3 residual GCPs, GCP 0, endpoint gathers}} -> node / edge embedding GCPs.
Edge features (e, xi) and frames depend only on the inputs, so their gradients stop at the embedding parameters.
"""
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

import gcpnet_oracle as O


def _dact(name: Optional[str], z: torch.Tensor) -> torch.Tensor:
    """d act(z) / dz for act in {silu, identity}."""
    if name != "silu":
        return torch.ones_like(z)
    s = torch.sigmoid(z)
    return s * (1 + z * (1 - s))


def _row_mean_frames(ei: torch.Tensor, frames: torch.Tensor, n: int) -> torch.Tensor:
    """fbar[n] = mean over the edges of row n of frames[e]  (node-side scalarize commutes with the row mean)."""
    tot = torch.zeros((n, 3, 3), dtype=frames.dtype).index_add_(0, ei[0], frames)
    cnt = torch.zeros(n, dtype=frames.dtype).index_add_(0, ei[0], torch.ones(ei.shape[1], dtype=frames.dtype))
    return tot / cnt.clamp(min=1)[:, None, None]


def gcp2_forward(sd, p, s, v, ei, frames, node_inputs, nonlin, feedforward_out=False, vector_out=True):
    """Same values as gcpnet_oracle.gcp2, plus the tape the backward needs."""
    W = lambda k: sd[p + k]
    vt = v.transpose(-1, -2)
    hid = vt @ W("vector_down.weight").t()                       # [M,3,H]
    n2 = (hid * hid).sum(dim=-2) + 1e-8                          # [M,H]
    vnorm = torch.sqrt(n2) + 1e-8
    vdf = (vt @ W("vector_down_frames.weight").t()).transpose(-1, -2)    # [M,3ch,3xyz]
    fr = _row_mean_frames(ei, frames, s.shape[0]) if node_inputs else frames
    q = torch.einsum("eax,ecx->eca", fr, vdf).reshape(s.shape[0], 9)
    merged = torch.cat((s, vnorm, q), dim=-1)
    tape = dict(p=p, v=v, hid=hid, n2=n2, vdf=vdf, fr=fr, merged=merged, s_in=s.shape[1], nonlin=nonlin,
                ff=feedforward_out, vector_out=vector_out)
    if feedforward_out:
        z1 = F.linear(merged, W("scalar_out.0.weight"), W("scalar_out.0.bias"))
        z = F.linear(F.silu(z1), W("scalar_out.2.weight"), W("scalar_out.2.bias"))
        tape["z1"] = z1
    else:
        z = F.linear(merged, W("scalar_out.weight"), W("scalar_out.bias"))
    tape["z"] = z
    s_out = O._act(nonlin[0], z)
    if not vector_out:
        return s_out, None, tape
    up = (hid @ W("vector_up.weight").t()).transpose(-1, -2)     # [M,Vout,3]
    gate = F.linear(O._act(nonlin[1], z), W("vector_out_scale.weight"), W("vector_out_scale.bias"))
    sg = torch.sigmoid(gate)
    tape.update(up=up, sg=sg)
    return s_out, up * sg.unsqueeze(-1), tape


def _acc(grads: Dict[str, torch.Tensor], key: str, val: torch.Tensor):
    grads[key] = grads[key] + val if key in grads else val


def gcp2_backward(sd, tape, ds_out, dv_out, grads) -> Tuple[torch.Tensor, torch.Tensor]:
    """(d s_out, d v_out) -> (d s, d v); parameter gradients accumulated into `grads`."""
    p = tape["p"]
    W = lambda k: sd[p + k]
    z, hid, v = tape["z"], tape["hid"], tape["v"]
    nonlin = tape["nonlin"]
    dz = ds_out * _dact(nonlin[0], z)
    dhid = torch.zeros_like(hid)
    if tape["vector_out"]:
        up, sg = tape["up"], tape["sg"]
        dup = dv_out * sg.unsqueeze(-1)                          # [M,Vout,3]
        dgate = (dv_out * up).sum(-1) * sg * (1 - sg)            # [M,Vout]
        a1 = O._act(nonlin[1], z)
        _acc(grads, p + "vector_out_scale.weight", dgate.t() @ a1)
        _acc(grads, p + "vector_out_scale.bias", dgate.sum(0))
        dz = dz + (dgate @ W("vector_out_scale.weight")) * _dact(nonlin[1], z)
        # up[m,o,x] = sum_h Wu[o,h] hid[m,x,h]
        _acc(grads, p + "vector_up.weight", torch.einsum("mox,mxh->oh", dup, hid))
        dhid = dhid + torch.einsum("mox,oh->mxh", dup, W("vector_up.weight"))
    merged = tape["merged"]
    if tape["ff"]:
        z1 = tape["z1"]
        a = F.silu(z1)
        _acc(grads, p + "scalar_out.2.weight", dz.t() @ a)
        _acc(grads, p + "scalar_out.2.bias", dz.sum(0))
        dz1 = (dz @ W("scalar_out.2.weight")) * _dact("silu", z1)
        _acc(grads, p + "scalar_out.0.weight", dz1.t() @ merged)
        _acc(grads, p + "scalar_out.0.bias", dz1.sum(0))
        dmerged = dz1 @ W("scalar_out.0.weight")
    else:
        _acc(grads, p + "scalar_out.weight", dz.t() @ merged)
        _acc(grads, p + "scalar_out.bias", dz.sum(0))
        dmerged = dz @ W("scalar_out.weight")
    s_in = tape["s_in"]
    h = hid.shape[-1]
    ds = dmerged[:, :s_in]
    dvnorm = dmerged[:, s_in:s_in + h]
    dq = dmerged[:, s_in + h:].reshape(-1, 3, 3)                 # [M, ch, axis]
    # vnorm = sqrt(n2) + 1e-8, n2 = sum_x hid^2 + 1e-8
    dhid = dhid + (dvnorm / torch.sqrt(tape["n2"])).unsqueeze(-2) * hid
    # q[m,c,a] = sum_x fr[m,a,x] vdf[m,c,x]
    dvdf = torch.einsum("mca,max->mcx", dq, tape["fr"])
    # vdf[m,c,x] = sum_i Wf[c,i] v[m,i,x] ; hid[m,x,h] = sum_i Wd[h,i] v[m,i,x]
    _acc(grads, p + "vector_down_frames.weight", torch.einsum("mcx,mix->ci", dvdf, v))
    _acc(grads, p + "vector_down.weight", torch.einsum("mxh,mix->hi", dhid, v))
    dv = torch.einsum("mcx,ci->mix", dvdf, W("vector_down_frames.weight")) + \
        torch.einsum("mxh,hi->mix", dhid, W("vector_down.weight"))
    return ds, dv


# ------------------------------------------------------------------------------------------------ message passing
def message_passing_forward(sd, p, cfg, h, chi, e, xi, ei, frames):
    r, c = ei[0], ei[1]
    ms = torch.cat((h[r], e, h[c]), dim=-1)
    mv = torch.cat((chi[r], xi, chi[c]), dim=1)
    act = ("silu", "silu")
    tapes = []
    s, v, t0 = gcp2_forward(sd, p + "message_fusion.0.", ms, mv, ei, frames, False, act)
    tapes.append(t0)
    for k in range(1, cfg.num_message_layers):
        ds, dv, tk = gcp2_forward(sd, p + f"message_fusion.{k}.", s, v, ei, frames, False, act)
        tapes.append(tk)
        s, v = s + ds, v + dv
    wa, ba = sd[p + "scalar_message_attention.0.weight"], sd[p + "scalar_message_attention.0.bias"]
    attn = torch.sigmoid(F.linear(s, wa, ba))
    flat = torch.cat((s * attn, v.reshape(v.shape[0], -1)), dim=-1)
    agg = torch.zeros((h.shape[0], flat.shape[1]), dtype=flat.dtype).index_add_(0, r, flat)
    vd = cfg.chi_hidden
    tape = dict(p=p, tapes=tapes, s=s, attn=attn, n=h.shape[0], hd=h.shape[1], vd=chi.shape[1], ed=e.shape[1],
                xd=xi.shape[1])
    return agg[:, :-3 * vd], agg[:, -3 * vd:].reshape(-1, vd, 3), tape


def message_passing_backward(sd, cfg, tape, ei, da_s, da_v, grads):
    """(d agg_s [N,256], d agg_v [N,32,3]) -> (dh [N,256], dchi [N,32,3], de [E,Ed], dxi [E,Xd,3])."""
    p = tape["p"]
    r, c = ei[0], ei[1]
    s, attn = tape["s"], tape["attn"]
    dsa = da_s[r]                                                # gradient of the gated scalar message
    dv = da_v[r]
    wa = sd[p + "scalar_message_attention.0.weight"]
    dpre = (dsa * s).sum(-1, keepdim=True) * attn * (1 - attn)   # [E,1]
    _acc(grads, p + "scalar_message_attention.0.weight", dpre.t() @ s)
    _acc(grads, p + "scalar_message_attention.0.bias", dpre.sum(0))
    ds = dsa * attn + dpre @ wa
    for k in range(cfg.num_message_layers - 1, 0, -1):          # residual GCPs: gradient flows to both branches
        din_s, din_v = gcp2_backward(sd, tape["tapes"][k], ds, dv, grads)
        ds, dv = ds + din_s, dv + din_v
    dms, dmv = gcp2_backward(sd, tape["tapes"][0], ds, dv, grads)
    hd, vd, ed, xd, n = tape["hd"], tape["vd"], tape["ed"], tape["xd"], tape["n"]
    dh = torch.zeros((n, hd), dtype=dms.dtype).index_add_(0, r, dms[:, :hd]).index_add_(0, c, dms[:, hd + ed:])
    dchi = torch.zeros((n, vd, 3), dtype=dms.dtype).index_add_(0, r, dmv[:, :vd]).index_add_(0, c, dmv[:, vd + xd:])
    return dh, dchi, dms[:, hd:hd + ed], dmv[:, vd:vd + xd]


# --------------------------------------------------------------------------------------------- interaction layer
def interaction_forward(sd, p, cfg, h, chi, e, xi, ei, frames, mask_f, x):
    a_s, a_v, tmp = message_passing_forward(sd, p + "interaction.", cfg, h, chi, e, xi, ei, frames)
    fs = torch.cat((a_s, h), dim=-1)
    fv = torch.cat((a_v, chi), dim=1)
    r_s, r_v, tff = gcp2_forward(sd, p + "feedforward_network.0.", fs, fv, ei, frames, True, (None, None),
                                 feedforward_out=True)
    h2 = (h + r_s) * mask_f[:, None]
    chi2 = (chi + r_v) * mask_f[:, None, None]
    _, pv, tpos = gcp2_forward(sd, p + "node_position_update_gcp.", h2, chi2, ei, frames, True, ("silu", "silu"))
    x2 = (x + pv[:, 0, :]) * mask_f[:, None]
    return h2, chi2, x2, dict(mp=tmp, ff=tff, pos=tpos, hd=h.shape[1], vd=chi.shape[1])


def interaction_backward(sd, cfg, tape, ei, mask_f, dh2, dchi2, dx2, grads):
    """(d h_out, d chi_out, d x_out) -> (d h_in, d chi_in, d x_in, d e, d xi)."""
    dx = dx2 * mask_f[:, None]
    dpv = torch.zeros((dx.shape[0], 1, 3), dtype=dx.dtype)
    dpv[:, 0, :] = dx
    zs = torch.zeros_like(tape["pos"]["z"])
    gh, gchi = gcp2_backward(sd, tape["pos"], zs, dpv, grads)    # position GCP: only its vector output is used
    dh2 = (dh2 + gh) * mask_f[:, None]
    dchi2 = (dchi2 + gchi) * mask_f[:, None, None]
    dfs, dfv = gcp2_backward(sd, tape["ff"], dh2, dchi2, grads)  # residual: d r_s = d h2, d r_v = d chi2
    hd, vd = tape["hd"], tape["vd"]
    da_s, dh = dfs[:, :hd], dh2 + dfs[:, hd:]
    da_v, dchi = dfv[:, :vd], dchi2 + dfv[:, vd:]
    mh, mchi, de, dxi = message_passing_backward(sd, cfg, tape["mp"], ei, da_s, da_v, grads)
    return dh + mh, dchi + mchi, dx, de, dxi


# ------------------------------------------------------------------------------------------------------ denoiser
def denoiser_forward_with_tape(sd, cfg, batch_index, mask, xh, t, context=None):
    dtype = torch.float32
    mask_f = mask.to(dtype)
    xh = xh.to(dtype) * mask_f[:, None]
    x_init, h_in = xh[:, :3], xh[:, 3:]
    ei = O.fully_connected_edge_index(batch_index, mask)
    chi_in = O.orientations(x_init)
    e_in, xi_in = O.edge_features(x_init, ei)
    h_in = torch.cat((h_in, t.to(dtype).reshape(-1, 1)), dim=-1)
    if cfg.num_context:
        h_in = torch.cat((h_in, context.to(dtype).reshape(xh.shape[0], cfg.num_context)), dim=-1)
    nmol = int(batch_index.max().item()) + 1
    _, x = O.centralize(x_init, batch_index, mask, nmol)
    frames = O.localize(x, ei)
    e, xi, te = gcp2_forward(sd, "gcp_embedding.edge_embedding.", e_in, xi_in, ei, frames, False, ("silu", "silu"))
    h, chi, tn = gcp2_forward(sd, "gcp_embedding.node_embedding.", h_in, chi_in, ei, frames, True, (None, None))
    layers = []
    for l in range(cfg.num_layers):
        h, chi, x, tl = interaction_forward(sd, f"interaction_layers.{l}.", cfg, h, chi, e, xi, ei, frames, mask_f, x)
        layers.append(tl)
    hp, _, tp = gcp2_forward(sd, "scalar_node_projection_gcp.", h, chi, ei, frames, True, (None, None),
                             vector_out=False)
    vel = (x - x_init) * mask_f[:, None]
    _, vel = O.centralize(vel, batch_index, mask, nmol)
    out = torch.cat((vel, hp[:, :cfg.num_h]), dim=-1)
    tape = dict(ei=ei, mask_f=mask_f, nmol=nmol, te=te, tn=tn, layers=layers, tp=tp, hp_cols=hp.shape[1])
    return out, tape


def denoiser_backward(sd, cfg, batch_index, mask, tape, d_out) -> Dict[str, torch.Tensor]:
    """Gradient of sum(net_out * d_out) with respect to every parameter in `sd`."""
    grads: Dict[str, torch.Tensor] = {}
    ei, mask_f, nmol = tape["ei"], tape["mask_f"], tape["nmol"]
    n = d_out.shape[0]
    # vel_c[i] = vel[i] - m_i * (sum_mol vel) / cnt ;  vel = (x_L - x_init) * m
    dvc = d_out[:, :3]
    cnt = torch.zeros(nmol, dtype=dvc.dtype).index_add_(0, batch_index, mask_f)
    back = torch.zeros((nmol, 3), dtype=dvc.dtype).index_add_(0, batch_index, dvc * mask_f[:, None]) / cnt[:, None]
    dx = (dvc - back[batch_index]) * mask_f[:, None]
    dhp = torch.zeros((n, tape["hp_cols"]), dtype=d_out.dtype)
    dhp[:, :cfg.num_h] = d_out[:, 3:]
    dh, dchi = gcp2_backward(sd, tape["tp"], dhp, None, grads)
    de_tot, dxi_tot = None, None
    for l in range(cfg.num_layers - 1, -1, -1):
        dh, dchi, dx, de, dxi = interaction_backward(sd, cfg, tape["layers"][l], ei, mask_f, dh, dchi, dx, grads)
        de_tot = de if de_tot is None else de_tot + de
        dxi_tot = dxi if dxi_tot is None else dxi_tot + dxi
    gcp2_backward(sd, tape["tn"], dh, dchi, grads)               # node embedding: inputs are data
    gcp2_backward(sd, tape["te"], de_tot, dxi_tot, grads)        # edge embedding
    return grads
