"""Golden fixtures of the RoBERTa text tower, from transformers' own RobertaModel.

    python tests/golden/make_golden_text.py        # writes tests/golden/text_*.npz   (build container only)

The reference builds `RobertaModel.from_pretrained('roberta-base')` (vgqa/core/language/bert.py:49) — no pretrained weights
offline, so a `RobertaModel(RobertaConfig(...roberta-base sizes...))` gets the deterministic synthetic weights
`O.synth_state_dict(seed, front_end_ch=..., text_tower=(layers, vocab))`; its `last_hidden_state` on synthetic token ids and the
reference `FeatureResizer` output (bert.py:69-73) are stored.  transformers version recorded in the file."""
import os
import sys

import numpy as np
import torch
import transformers
from transformers import RobertaConfig, RobertaModel

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
from oracle import vgqa_oracle as O  # noqa: E402
from ref_loader import load_feature_resizer  # noqa: E402

# name, seed, B, L, layers, vocab, pad_tail
CASES = [("text_tiny_L5_2layers", 0, 2, 5, 2, 300, 2),
         ("text_base_L20_12layers", 1, 3, 20, 12, 2000, 4),
         ("text_base_L64_12layers", 2, 2, 64, 12, 2000, 0)]

if __name__ == "__main__":
    FR = load_feature_resizer()
    for name, seed, B, L, layers, vocab, pad_tail in CASES:
        ch = (128, 64, 768)
        sd = O.synth_state_dict(seed, front_end_ch=ch, text_tower=(layers, vocab))
        cfg = RobertaConfig(vocab_size=vocab, max_position_embeddings=514, type_vocab_size=1, pad_token_id=1, layer_norm_eps=1e-5,
                            num_hidden_layers=layers, hidden_size=768, num_attention_heads=12, intermediate_size=3072)
        m = RobertaModel(cfg, add_pooling_layer=False).eval()
        body = {k[len("text_encoder.body."):]: torch.from_numpy(v) for k, v in sd.items() if k.startswith("text_encoder.body.")}
        missing, unexpected = m.load_state_dict(body, strict=False)
        assert not unexpected and all("position_ids" in k or "token_type_ids" in k for k in missing), (missing, unexpected)
        rs = FR(input_feat_size=768, output_feat_size=256, dropout=0.1).eval()
        rs.load_state_dict({k[len("text_encoder.resizer."):]: torch.from_numpy(v) for k, v in sd.items()
                            if k.startswith("text_encoder.resizer.")})
        ids, pad = O.synth_text_ids(seed, B, L, vocab, pad_tail)
        with torch.no_grad():
            out = m(input_ids=torch.from_numpy(ids).long(), attention_mask=torch.from_numpy(~pad).long())
            hid = out.last_hidden_state
            text = rs(hid.transpose(0, 1))                      # (L, B, 256), bert.py:69,73
            # how far torch's OWN bf16 run (CPU autocast) of the same modules lands from fp32: the yardstick for the tower's
            # intermediate tensors (the north-star 2e-2 bar applies to the path's final boxes / logits)
            with torch.autocast("cpu", dtype=torch.bfloat16):
                hid16 = m(input_ids=torch.from_numpy(ids).long(), attention_mask=torch.from_numpy(~pad).long()).last_hidden_state
                text16 = rs(hid16.transpose(0, 1))
        keep_t = torch.from_numpy(~pad)
        ac_h = float((hid16.float() - hid).abs()[keep_t].max())
        ac_t = float((text16.float() - text).abs()[keep_t.T].max())
        np.savez_compressed(os.path.join(HERE, name + ".npz"), seed=seed, B=B, L=L, layers=layers, vocab=vocab, pad_tail=pad_tail,
                            front_end_ch=np.asarray(ch), last_hidden_state=hid.numpy(), text_resized=text.numpy(),
                            bf16_autocast_err_hidden=np.float32(ac_h), bf16_autocast_err_text=np.float32(ac_t),
                            transformers_version=transformers.__version__, torch_version=torch.__version__)
        mine = O.roberta_encoder(sd, ids, pad)
        keep = ~pad
        print(name, "oracle vs transformers max-abs (unpadded rows):", float(np.abs(mine - hid.numpy())[keep].max()),
              " |hidden| mean", float(np.abs(hid.numpy()).mean()), " torch bf16 autocast err hidden/text:", ac_h, ac_t)
