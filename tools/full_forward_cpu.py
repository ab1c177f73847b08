"""BASELINE.json configs[0]: the reference's FULL `VSTGNet.forward` on the host cores — configs/grounding_vidstg_mini.yaml (224 px, 32
frames), synthetic clip, random-init weights (ResNet101 + Video-Swin-T + RoBERTa-base + the hot path, two decoder passes).

    python tools/full_forward_cpu.py            # prints one JSON line: seconds per forward, per-part times

Test infrastructure / reported baseline only (bench.py --impl reference calls it): the reference's own modules are imported from
/root/reference where it is mounted, else from the byte-compiled oracle/_ref (tools/make_oracle_ref.py).  What the reference
needs and this image lacks is replaced by stand-ins, exactly the list of SURVEY.md §8c: `yacs` (CfgNode), `easydict`, `timm`
(DropPath / trunc_normal_), `torchtext`, `pytorch_pretrained_bert`, `decord`, `ffmpeg` (import-time only); torchvision's
`pretrained=True` download → random init; RoBERTa `from_pretrained` (no weights / vocabulary offline) → a random-init roberta-base
fed seeded token ids (the tokenizer is a string → ids stand-in); `annos/test.json` absent → `verb_label2` set as
vgqa/inference/grounding.py:123-126 does.
"""
import copy
import importlib
import importlib.machinery
import json
import os
import sys
import time
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))


class CfgNode(dict):
    """Attribute-access dict tree: the subset of yacs.config.CfgNode the reference uses (clone / merge_from_list / freeze)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v

    def clone(self):
        return copy.deepcopy(self)

    def merge_from_list(self, kv):
        for key, val in zip(kv[0::2], kv[1::2]):
            node = self
            parts = key.split(".")
            for p in parts[:-1]:
                node = node[p]
            node[parts[-1]] = val

    def freeze(self):
        pass

    def defrost(self):
        pass


def install_stand_ins():
    import torch
    import transformers                              # before the stand-ins: it probes optional packages (timm, ...) at import
    import ref_loader
    ref_loader._install_shims()                      # easydict, tqdm
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__path__ = []
        m.__spec__ = importlib.machinery.ModuleSpec(name, None)   # transformers probes optional packages with find_spec
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules.setdefault(name, m)
        return sys.modules[name]
    mod("yacs"); mod("yacs.config", CfgNode=CfgNode)
    sys.modules["yacs"].config = sys.modules["yacs.config"]

    class DropPath(torch.nn.Module):
        def __init__(self, p=0.0):
            super().__init__()

        def forward(self, x):
            return x

    mod("timm"); mod("timm.models"); mod("timm.models.layers", DropPath=DropPath, trunc_normal_=torch.nn.init.trunc_normal_)
    mod("torchtext"); mod("decord"); mod("ffmpeg")
    mod("pytorch_pretrained_bert"); mod("pytorch_pretrained_bert.modeling", BertModel=object)
    mod("pytorch_pretrained_bert.tokenization", BertTokenizer=object)
    # torchvision: pretrained=True would download → random init
    import torchvision
    if not getattr(torchvision.models, "_vgqa_patched", False):
        for name in ("resnet50", "resnet101"):
            orig = getattr(torchvision.models, name)
            setattr(torchvision.models, name, (lambda o: lambda pretrained=False, **kw: o(weights=None, **kw))(orig))
        torchvision.models._vgqa_patched = True
    # RoBERTa: random-init roberta-base + a deterministic string → ids stand-in for the tokenizer
    import transformers

    class StubTokenizer:
        def batch_encode_plus(self, texts, padding="longest", return_tensors="pt"):
            import zlib
            rows = [[0] + [3 + zlib.crc32(w.encode()) % 50000 for w in t.split()] + [2] for t in texts]
            n = max(len(r) for r in rows)
            ids = torch.tensor([r + [1] * (n - len(r)) for r in rows])
            att = torch.tensor([[1] * len(r) + [0] * (n - len(r)) for r in rows])
            return transformers.BatchEncoding({"input_ids": ids, "attention_mask": att})

    cfg_rb = transformers.RobertaConfig(vocab_size=50265, max_position_embeddings=514, type_vocab_size=1, pad_token_id=1)
    transformers.RobertaModel.from_pretrained = classmethod(lambda cls, name, *a, **k: transformers.RobertaModel(cfg_rb))
    transformers.RobertaTokenizerFast.from_pretrained = classmethod(lambda cls, *a, **k: StubTokenizer())


def build_reference_model(frames=32):
    import torch
    import ref_loader
    install_stand_ins()
    if not ref_loader.reference_modules_available():
        raise RuntimeError("reference modules not found (tools/make_oracle_ref.py builds oracle/_ref where /root/reference is mounted)")
    if ref_loader.REF_ROOT == ref_loader._COMPILED and not any(isinstance(f, ref_loader._CompiledFinder) for f in sys.meta_path):
        sys.meta_path.insert(0, ref_loader._CompiledFinder(ref_loader.REF_ROOT))
    v = os.path.join(ref_loader.REF_ROOT, "vgqa")
    ref_loader._namespace_pkg("vgqa", v)              # the top-level __init__ pulls in data / inference / training: bypassed
    for sub in ("utils", "training"):
        ref_loader._namespace_pkg("vgqa." + sub, os.path.join(v, sub))
    cfg = importlib.import_module("vgqa.config").cfg.clone()
    cfg.merge_from_list(["INPUT.RESOLUTION", 224, "INPUT.TRAIN_SAMPLE_NUM", frames, "DATA_DIR", "data/vidstg"])   # grounding_vidstg_mini.yaml
    torch.manual_seed(0)
    VSTGNet = importlib.import_module("vgqa.core.grounding_net").VSTGNet
    model = VSTGNet(cfg).eval()
    model.verb_label2 = {"0": {"sub": "", "verb_index_list": [], "adj_index_list": []}}          # grounding.py:123-126
    NestedTensor = importlib.import_module("vgqa.utils.training_utils").NestedTensor
    return model, NestedTensor


def full_forward_seconds(frames=32, res=224, repeats=1):
    import torch
    torch.set_num_threads(os.cpu_count())
    model, NestedTensor = build_reference_model(frames)
    g = torch.Generator().manual_seed(1)
    videos = NestedTensor(torch.randn(frames, 3, res, res, generator=g), torch.zeros(frames, res, res, dtype=torch.bool), [frames])
    targets = [{"item_id": 0, "actioness": torch.ones(frames)}]
    parts = {}
    hooks = []

    def timed(name, m):
        def pre(mod, args):
            parts["_t_" + name] = time.perf_counter()

        def post(mod, args, out):
            parts[name] = parts.get(name, 0.0) + time.perf_counter() - parts["_t_" + name]
        hooks.append(m.register_forward_pre_hook(pre)); hooks.append(m.register_forward_hook(post))

    for name in ("vis_encoder", "vid", "text_encoder", "ground_encoder", "ground_decoder"):
        timed(name, getattr(model, name))
    times = []
    with torch.no_grad():
        for i in range(1 + repeats):          # first call = warm-up (allocator, thread pools)
            for k in list(parts):
                parts.pop(k)
            t0 = time.perf_counter()
            out = model(videos, ["a person jumping over the fence"], targets)
            times.append(time.perf_counter() - t0)
    for h in hooks:
        h.remove()
    best = min(times[1:]) if len(times) > 1 else times[0]
    return {"seconds_per_forward": best, "first_call_seconds": times[0], "frames": frames, "resolution": res,
            "parts_seconds_last_call": {k: round(v, 3) for k, v in parts.items() if not k.startswith("_t_")},
            "threads": torch.get_num_threads(), "pred_boxes_shape": list(out["pred_boxes"].shape),
            "what": "the reference's full VSTGNet.forward (ResNet101 + Video-Swin-T + RoBERTa-base + hot path, two decoder passes), "
                    "configs/grounding_vidstg_mini.yaml, synthetic clip, random-init weights, fp32, all host threads"}


if __name__ == "__main__":
    print(json.dumps(full_forward_seconds()))
