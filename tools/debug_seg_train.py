"""Per-module isolation of the segmentation network's training tape: every MiT block / the decode head is run with the
ORACLE's exact input and the ORACLE's exact output gradient, so that an error is attributable to one module instead of
being accumulated noise.  Diagnostic tool (GPU box only); prints one line per comparison."""
import copy
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

from gpu_checks import DEV, rel_err  # noqa: E402
from oracle import segmif_oracle as O  # noqa: E402
from segmif_b200 import ops, synth  # noqa: E402
from segmif_b200.core import seg_train as T  # noqa: E402
from segmif_b200.core.model_fusion import Network3  # noqa: E402

BF16, F32 = torch.bfloat16, torch.float32


def main(droppath_on=False):
    B, H, W = 2, 64, 96
    backbone = "mit_b1"
    net0 = synth.load_synthetic(Network3(backbone, 9, 256, None), 0)
    sd = {k: v.clone() for k, v in net0.state_dict().items()}
    names = [k for k, _ in net0.named_parameters() if not k.endswith("classifier.weight")]
    gen = torch.Generator().manual_seed(21)
    x = torch.rand(B, 3, H, W, generator=gen)
    drop = ((torch.rand(B, 256, generator=gen) >= 0.1).float() / 0.9)
    h, w = H // 4, W // 4
    cot = torch.randn(B, 9, h, w, generator=gen) / (B * h * w)
    dps = [((torch.rand(B, generator=gen) < 0.8).float() / 0.8, (torch.rand(B, generator=gen) < 0.8).float() / 0.8) for _ in range(8)]
    droppath = dps if droppath_on else None

    # ---- oracle with every block boundary retained
    leaves = {k: sd[k].clone().requires_grad_(True) for k in names}
    full = dict(sd)
    full.update(leaves)
    esd = O._sub(full, "denoise_net.encoder")
    mean = torch.tensor(O.IMAGENET_MEAN).view(1, 3, 1, 1)
    std = torch.tensor(O.IMAGENET_STD).view(1, 3, 1, 1)
    cur = (x * 255 - mean) / std
    cfg = O.MIT_CONFIGS[backbone]
    toks, feats = [], []            # toks[s] = [tok after patch embed, after block 0, ...]
    bi = 0
    for s in range(4):
        patch, stride = (7, 4) if s == 0 else (3, 2)
        tok, hh, ww = O.overlap_patch_embed(cur, esd, f"patch_embed{s + 1}", patch, stride)
        tok.retain_grad()
        lst = [tok]
        for i in range(cfg["depths"][s]):
            tok = O.mit_block(tok, hh, ww, esd, f"block{s + 1}.{i}", O.MIT_HEADS[s], O.MIT_SR[s], None if droppath is None else droppath[bi])
            tok.retain_grad()
            lst.append(tok)
            bi += 1
        toks.append((lst, hh, ww))
        tokn = O.layer_norm(tok, esd, f"norm{s + 1}", O.BLOCK_LN_EPS)
        cur = tokn.reshape(B, hh, ww, -1).permute(0, 3, 1, 2).contiguous()
        cur.retain_grad()
        feats.append(cur)
    lg_ref = O.segformer_head(feats, O._sub(full, "denoise_net.decoder"), train_bn=True, dropout_scale=drop)
    (lg_ref * cot).sum().backward()

    net = copy.deepcopy(net0).to(DEV).train()
    enc, head = net.denoise_net.encoder, net.denoise_net.decoder
    lookup = dict(net.named_parameters())

    def newg():
        return {n: torch.zeros(lookup[n].shape, dtype=F32, device=DEV) for n in names}

    def report(tag, g, prefix):
        worst, wn = 0.0, ""
        for n in names:
            if n.startswith(prefix):
                e = rel_err(g[n], leaves[n].grad)
                if e > worst:
                    worst, wn = e, n
                if e > 0.03:
                    print(f"    {n:70s} err={e:.3e} |ref|max={float(leaves[n].grad.abs().max()):.3e}")
        print(f"  {tag}: worst param grad err {worst:.3e} ({wn})")

    # ---- blocks in isolation
    bi = 0
    for s in range(4):
        lst, hh, ww = toks[s]
        N = hh * ww
        for i, blk in enumerate(getattr(enc, f"block{s + 1}")):
            xin = lst[i].detach().reshape(B * N, -1).contiguous().to(DEV)
            inj = droppath[bi] if droppath is not None else None
            s1 = inj[0].to(DEV) if inj is not None else None
            s2 = inj[1].to(DEV) if inj is not None else None
            x3, sv = T._block_forward(blk, xin, B, N, hh, ww, (s1, s2))
            e_f = rel_err(x3.reshape(B, N, -1), lst[i + 1])
            dx = lst[i + 1].grad.reshape(B * N, -1).contiguous().to(DEV).clone()
            g = newg()
            T._block_backward(blk, sv, dx, B, N, g, f"denoise_net.encoder.block{s + 1}.{i}.")
            e_b = rel_err(dx.reshape(B, N, -1), lst[i].grad)
            print(f"block{s + 1}.{i}: fwd err {e_f:.3e}  dx err {e_b:.3e}")
            report(f"block{s + 1}.{i}", g, f"denoise_net.encoder.block{s + 1}.{i}.")
            bi += 1

    # ---- head in isolation
    stages = []
    for f in feats:
        t = f.detach().permute(0, 2, 3, 1).reshape(-1, f.shape[1]).contiguous().to(BF16).to(DEV)
        stages.append((t, f.shape[2], f.shape[3]))
    logits, htape = T.head_forward(head, stages, B, {"dropout2d": drop.to(DEV)})
    print(f"head: logits err {rel_err(logits.permute(0, 3, 1, 2), lg_ref):.3e}")
    g = newg()
    douts = T.head_backward(head, htape, cot.permute(0, 2, 3, 1).contiguous().to(DEV), B, g, "denoise_net.decoder.")
    for s in range(4):
        f = feats[s]
        print(f"  head d feat{s + 1} err {rel_err(douts[s].float().reshape(B, f.shape[2], f.shape[3], -1).permute(0, 3, 1, 2), f.grad):.3e}")
    report("head", g, "denoise_net.decoder.")

    # ---- whole encoder with the oracle's feature gradients (accumulated noise, for comparison)
    sc, sh = net._input_affine(torch.device(DEV))
    masks = {}
    if droppath is not None:
        for i, (a, b) in enumerate(droppath):
            masks[("droppath", i)] = (a.to(DEV), b.to(DEV))
    outs, etape = T.encoder_forward(enc, x.to(DEV), sc, sh, masks)
    for s in range(4):
        f = feats[s]
        print(f"encoder stage{s + 1} out err {rel_err(outs[s][0].float().reshape(B, f.shape[2], f.shape[3], -1).permute(0, 3, 1, 2), f):.3e}")
        lst = toks[s][0]
        st = etape["stages"][s]
        print(f"   patch-embed tokens err {rel_err(st['blocks'][0]['x'].reshape(B, -1, lst[0].shape[-1]), lst[0]):.3e}")
        for i, sv in enumerate(st["blocks"]):
            nxt = st["blocks"][i + 1]["x"] if i + 1 < len(st["blocks"]) else st["tok_final"]
            print(f"   after block {i}: err {rel_err(nxt.reshape(B, -1, lst[0].shape[-1]), lst[i + 1]):.3e}  x2-vs-x err n/a")
    g = newg()
    dfe = [f.grad.permute(0, 2, 3, 1).reshape(-1, f.shape[1]).contiguous().to(BF16).to(DEV) for f in feats]
    # the oracle's feature gradients already contain the carry from the next stage's patch embedding: remove it by
    # feeding only the head's share -- recomputed from the head in isolation above
    T.encoder_backward(enc, etape, [d.clone() for d in douts], g, "denoise_net.encoder.", False)
    report("encoder(full chain, head grads from isolated head)", g, "denoise_net.encoder.")


if __name__ == "__main__":
    main(droppath_on=len(sys.argv) > 1 and sys.argv[1] == "droppath")
