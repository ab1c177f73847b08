"""Golden fixture of the reference's WHOLE `VSTGNet.forward` (build container only):

    python tests/golden/make_golden_full.py        # writes tests/golden/full_vstgnet_*.npz

The unmodified reference model (vgqa/core/grounding_net.py:24-203: ResNet101 + PositionEmbeddingSine, Video-Swin-T, RoBERTa-base +
FeatureResizer, input_proj / input_proj2, CrossModalEncoder, the classifiers, two decoder passes, the heads) is built by
tools/full_forward_cpu.py (which also lists the stand-ins for what this image lacks: yacs, timm, the RoBERTa download → random-init
roberta-base fed seeded token ids), loaded with the seeded synthetic weights of vgqa_b200/synth.py (hot path + front end + text
tower, ResNet101, Video-Swin-T — regenerated from the seed by the test, never stored) and run in fp32 on seeded frames.  Stored: the
token ids, every output of the forward, the two frame selections (and the first pass's actioness that decides the second)."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from oracle import vgqa_oracle as O  # noqa: E402

# (name, T, R, seed, sentence, (pad_right, pad_bottom)): padded pixels are zero and masked, as the reference's collate leaves them
CASES = [("full_vstgnet_T16_224_s0", 16, 224, 0, "a person jumping over the fence", (0, 0)),
         ("full_vstgnet_T8_224_masked_s1", 8, 224, 1, "the dog that runs behind the red car", (64, 32)),
         ("full_vstgnet_T12_256_s2", 12, 256, 2, "two children play with a ball", (0, 0))]      # 8x8 maps; Video-Swin pads 12 → 16 frames, 64 → 70 ...
FRONT_END_CH = (2048, 768, 768)
TEXT_TOWER = (12, 50265)


def full_frames(seed, T, R, pad=(0, 0)):
    rng = np.random.Generator(np.random.PCG64(17000 + seed))
    x = rng.standard_normal((T, 3, R, R), dtype=np.float32)
    x[:, :, :, R - pad[0]:] = 0.0
    x[:, :, R - pad[1]:, :] = 0.0
    return x


def full_mask(T, R, pad=(0, 0)):
    m = np.zeros((T, R, R), bool)
    m[:, :, R - pad[0]:] = True
    m[:, R - pad[1]:, :] = True
    return m


def full_state_dict(seed):
    sd = O.synth_state_dict(seed, front_end_ch=FRONT_END_CH, text_tower=TEXT_TOWER)
    sd.update(O.synth_resnet101(seed))
    sd.update(O.synth_swin_backbone(seed))
    return sd


if __name__ == "__main__":
    import full_forward_cpu as FF
    torch.set_num_threads(os.cpu_count())
    for name, T, R, seed, sentence, pad in CASES:
        model, NestedTensor = FF.build_reference_model(T)
        sd = full_state_dict(seed)
        missing, unexpected = model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=False)
        # `missing` = parameters the reference constructs but never reads in forward (temporal_layers, class / positional embeddings
        # of the classifiers, gf_mlp, time_fc ... — SURVEY.md §8a) plus buffers; they keep their torch init.  Nothing may be unexpected.
        print("missing:", len(missing), "unexpected:", len(unexpected))
        assert not unexpected
        taps = {"dec": [], "ids": None}
        h1 = model.ground_decoder.register_forward_hook(lambda m, a, out: taps["dec"].append(out[1].detach()))
        tok = model.text_encoder.tokenizer
        orig = tok.batch_encode_plus

        def spy(texts, **kw):
            enc = orig(texts, **kw)
            taps["ids"] = enc["input_ids"].clone()
            return enc
        tok.batch_encode_plus = spy
        x = full_frames(seed, T, R, pad)
        videos = NestedTensor(torch.from_numpy(x), torch.from_numpy(full_mask(T, R, pad)), [T])
        targets = [{"item_id": 0, "actioness": torch.ones(T)}]
        with torch.no_grad():
            out = model(videos, [sentence], targets)
            act1 = model.action_embed(taps["dec"][0])[-1].squeeze().sigmoid()
        h1.remove()
        att = out["att_sequences"][0]
        c1 = torch.nonzero(att > model.theta).flatten().tolist() or list(range(T))
        c2 = torch.nonzero(act1 > 0.5).flatten().tolist() or list(range(T))
        rec = dict(T=T, R=R, seed=seed, pad=np.asarray(pad, np.int32), sentence=sentence, torch_version=torch.__version__, text_ids=taps["ids"].numpy().astype(np.int32),
                   choose1=np.asarray(c1, np.int32), choose2=np.asarray(c2, np.int32), actioness_pass1=act1.numpy(),
                   theta_margin=np.float32((att - model.theta).abs().min()), act_margin=np.float32((act1 - 0.5).abs().min()))
        for k in ("pred_boxes", "pred_sted", "pred_actioness", "logits_f_m", "logits_f_a", "logits_r_a", "logits_r_m", "att_sequences"):
            rec[k] = out[k].numpy()
        for i, a in enumerate(out["aux_outputs"]):
            for k in ("pred_boxes", "pred_sted", "pred_actioness"):
                rec[f"aux{i}_{k}"] = a[k].numpy()
        print(name, "ids", rec["text_ids"].tolist(), "choose1", len(c1), "choose2", len(c2), "of", T, "margins", float(rec["theta_margin"]),
              float(rec["act_margin"]), "pred_boxes[0]", rec["pred_boxes"][0], "sted shape", rec["pred_sted"].shape)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **rec)
