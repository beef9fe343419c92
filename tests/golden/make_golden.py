"""Generate tests/golden/*.npz by running the UNMODIFIED reference (build container only).

    python tests/golden/make_golden.py

Imports /root/reference/{tdnn_layer,main}.py through oracle/ref_loader.py (third-party
imports stubbed), runs them in eval mode on seeded synthetic inputs, and stores
inputs-by-seed + outputs.  The weights are not stored (20 MB): they are the
reference model's own default initialisation under torch.manual_seed(0) and the
fixture records their SHA-256, which oracle.make_state_dict must reproduce.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_loader, xvector_oracle as ox  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    torch.set_num_threads(8)
    tdnn_layer, ref_main = ref_loader.load()

    # ---- 1. known-answer fixtures of the reference: extra/time_context_test.py + docstring example
    rows = [list(range(1, 16)), [4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 1, 2, 3],
            [7, 8, 9, 10, 11, 12, 13, 14, 15, 1, 2, 3, 4, 5, 6], [10, 11, 12, 13, 14, 15, 1, 2, 3, 4, 5, 6, 7, 8, 9],
            [13, 14, 15, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12]]
    x = torch.tensor(rows).unsqueeze(-1)
    kat = {"x": x.numpy()}
    for name, ctx in (("c5", [-2, -1, 0, 1, 2]), ("c2", [-2, 2]), ("c5d2", [-4, -2, 0, 2, 4]), ("c11", list(range(-5, 6)))):
        kat[name + "_ctx"] = np.asarray(ctx)
        kat[name] = torch.cat(tdnn_layer.get_time_context(x, ctx), 2).numpy()
    xd = torch.tensor([[[1, 2], [3, 4], [5, 6], [7, 8], [9, 0]]])
    kat["doc_x"] = xd.numpy()
    kat["doc"] = torch.cat(tdnn_layer.get_time_context(xd, [-1, 0, 1]), 2).numpy()
    np.savez_compressed(os.path.join(OUT, "time_context_kat.npz"), **kat)

    # ---- 2. the reference model under seed 0 (+ the oracle's BN randomisation copied in)
    torch.manual_seed(0)
    models = {6: ref_main.XVectorModel(x_vec_extract_layer=6).eval()}
    default_digest = ox.state_dict_digest({k: v for k, v in models[6].state_dict().items() if "accuracy" not in k})
    sd = ox.make_state_dict(seed=0, randomize_bn=True)
    missing = models[6].load_state_dict(sd, strict=False)
    assert not missing.unexpected_keys, missing
    models[7] = ref_main.XVectorModel(x_vec_extract_layer=7).eval()
    models[7].load_state_dict(sd, strict=False)
    models[3] = ref_main.XVectorModel(x_vec_extract_layer=3).eval()  # "anything else behaves as 6"
    models[3].load_state_dict(sd, strict=False)

    g = {"default_init_digest": default_digest, "state_digest": ox.state_dict_digest(sd)}
    with torch.no_grad():
        # fixed-length batch, the canonical (·,299,24) shape of main.py:113 and BASELINE's 300
        for tag, (b, t, seed) in {"b4_t299": (4, 299, 11), "b8_t300": (8, 300, 1234), "b3_t16": (3, 16, 5),
                                  "b2_t15": (2, 15, 6)}.items():
            xx = ox.synth_mfcc(b, t, seed=seed)
            g[tag + "_shape_seed"] = np.asarray([b, t, seed])
            g[tag + "_l6"] = models[6].extract_x_vec(xx).numpy()
            g[tag + "_l7"] = models[7].extract_x_vec(xx).numpy()
            g[tag + "_l3"] = models[3].extract_x_vec(xx).numpy()
            g[tag + "_fwd"] = models[6](xx).numpy()
        # per-layer activations + pooled statistics on a small batch
        xx = ox.synth_mfcc(2, 40, seed=21)
        h = xx
        for i, layer in enumerate(models[6].time_context_layers):
            h = layer(h)
            g[f"act_l{i + 1}"] = h.numpy()
        g["act_pool"] = models[6].stat_pool(h).numpy()
        # ragged: each utterance alone at its true length (the reference has no masking)
        lens = ox.synth_lengths(12, 16, 420, seed=2)
        utts = ox.synth_ragged(lens, seed=77)
        g["ragged_lengths"] = lens
        g["ragged_l6"] = np.stack([models[6].extract_x_vec(u[None])[0].numpy() for u in utts])
        # a standalone TdnnLayer without BN and one with an odd channel count
        torch.manual_seed(3)
        lay = tdnn_layer.TdnnLayer(input_size=40, output_size=96, context=[-3, 0, 3], batch_norm=False).eval()
        xl = ox.synth_mfcc(3, 50, 40, seed=9)
        g["layer_nobn_w"] = lay.linear.weight.detach().numpy()
        g["layer_nobn_b"] = lay.linear.bias.detach().numpy()
        g["layer_nobn_y"] = lay(xl).numpy()
    np.savez_compressed(os.path.join(OUT, "xvector_golden.npz"), **g)
    for f in ("time_context_kat.npz", "xvector_golden.npz"):
        print(f, os.path.getsize(os.path.join(OUT, f)), "bytes")


if __name__ == "__main__":
    main()
