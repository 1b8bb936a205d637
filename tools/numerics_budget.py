"""CPU emulation of the GPU numerics (16-bit MMA operands, fp32 accumulate / residual / LN / softmax) to see which
operand type meets |dscore| <= 1e-3 against the fp32 oracle before spending GPU time.  Dev tool, not product."""
import math
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from kddcup_2020_multimodalitiesrecall_2nd_place_b200 import synth
from kddcup_2020_multimodalitiesrecall_2nd_place_b200.config import ZK, LDS, ModelConfig
from oracle import imagebert as ob


def emu_forward(w, inp, cfg, dt, res32=True, kind=ZK):
    r = lambda x: x.to(dt).float()
    def gemm(x16, name_k, name_b):  # x16 already rounded
        return x16 @ r(w[name_k]) + w[name_b]
    ln = ob.layer_norm
    if kind == ZK:
        feat = torch.relu(r(inp["feats"]) @ r(w["kdd_conv2/weights"][0, 0]) + w["kdd_conv2/biases"])
        label = ob.zk_label_term(inp["label_ids"], w)   # tables: approx fp32
        box = inp["boxes"] @ w["kdd_dense1/weights"] + w["kdd_dense1/biases"]
        t = r(label + box + feat)
        region = t @ r(w["kdd_featureemb/fully_connected/weights"]) + w["kdd_featureemb/fully_connected/biases"]
        x = ob.zk_embeddings(inp["query_ids"], inp["segment_ids"], region, w)
        Lq = cfg.lq; R = cfg.nbox
        qmask = torch.arange(Lq)[None] < inp["len_query"].long()[:, None]
        bmask = torch.arange(R)[None] < inp["num_boxes"].long()[:, None]
        km = torch.cat([qmask, bmask], 1).float()
    else:
        region = r(inp["feats"]) @ r(w["featureemb/fully_connected/weights"]) + w["featureemb/fully_connected/biases"]
        label = ob.lds_label_term(inp["label_ids"], w)
        E = w["bert/embeddings/word_embeddings"]
        text = E[inp["query_ids"].long()] + w["bert/embeddings/token_type_embeddings"][inp["segment_ids"].long()] + w["bert/embeddings/position_embeddings"][:cfg.lq]
        text = ln(text, w["bert/embeddings/LayerNorm/gamma"], w["bert/embeddings/LayerNorm/beta"])
        x = torch.cat([text, region, label], 1)
        km = None
    B, S, H = x.shape
    x32 = x
    for i in range(cfg.n_layers):
        p = f"bert/encoder/layer_{i}"
        x16 = r(x32)
        xr = x32 if res32 else x16
        q = r(gemm(x16, p + "/attention/self/query/kernel", p + "/attention/self/query/bias")).view(B, S, 12, 64).transpose(1, 2)
        k = r(gemm(x16, p + "/attention/self/key/kernel", p + "/attention/self/key/bias")).view(B, S, 12, 64).transpose(1, 2)
        v = r(gemm(x16, p + "/attention/self/value/kernel", p + "/attention/self/value/bias")).view(B, S, 12, 64).transpose(1, 2)
        s = q @ k.transpose(-1, -2) * 0.125
        if km is not None:
            s = s + (1 - km)[:, None, None, :] * -10000.0
        e = torch.exp(s - s.max(-1, keepdim=True).values)
        ctx = (r(e) @ v) / e.sum(-1, keepdim=True)
        ctx = r(ctx.transpose(1, 2).reshape(B, S, H))
        y = gemm(ctx, p + "/attention/output/dense/kernel", p + "/attention/output/dense/bias") + xr
        a32 = ln(y, w[p + "/attention/output/LayerNorm/gamma"], w[p + "/attention/output/LayerNorm/beta"])
        a16 = r(a32)
        ar = a32 if res32 else a16
        h = r(ob.gelu_tanh(gemm(a16, p + "/intermediate/dense/kernel", p + "/intermediate/dense/bias")))
        z = gemm(h, p + "/output/dense/kernel", p + "/output/dense/bias") + ar
        x32 = ln(z, w[p + "/output/LayerNorm/gamma"], w[p + "/output/LayerNorm/beta"])
    cls = r(x32[:, 0])
    pooled = torch.tanh(cls @ r(w["bert/pooler/dense/kernel"]) + w["bert/pooler/dense/bias"])
    if kind == ZK:
        return ob.amsoftmax_probs(pooled, inp["labels"], w), pooled
    logits = pooled @ w["cls/seq_relationship/output_weights"].t() + w["cls/seq_relationship/output_bias"]
    return torch.softmax(logits, -1), pooled


if __name__ == "__main__":
    torch.set_num_threads(8)
    B = 48
    for kind in (ZK, LDS):
        for tl in (False, True):
            cfg = ModelConfig(kind, n_layers=12, lq=32, nbox=36, vocab=3000)
            w = ob.to_torch(synth.make_weights(cfg, seed=1, trained_like=tl))
            inp = ob.to_torch(synth.make_inputs(cfg, B, seed=1))
            with torch.no_grad():
                fwd = ob.zk_forward if kind == ZK else ob.lds_forward
                ref = fwd(w, inp, cfg.n_layers)
                for dt in (torch.bfloat16, torch.float16):
                    for res32 in (True, False):
                        pr, pooled = emu_forward(w, inp, cfg, dt, res32, kind)
                        d = (pr[:, 1] - ref["probs"][:, 1]).abs()
                        dp = (pooled - ref["pooled"]).abs().max()
                        print(f"{kind} trained_like={tl} {str(dt)[6:]} res32={res32}: max|dscore|={d.max():.2e} mean={d.mean():.2e} "
                              f"max|dpooled|={dp:.2e}  score range [{ref['probs'][:,1].min():.3f},{ref['probs'][:,1].max():.3f}]", flush=True)
