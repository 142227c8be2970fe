"""Host-side checks that need no GPU: the C-ABI library loads, exports every symbol the header declares (and
the ctypes table binds exactly those), fails loudly without a device, and the Python mirror keeps the
reference's state-dict schema."""
import ctypes as C
import os
import re
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "rgqa_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return set(re.findall(r"\b(rgqa_[a-z0-9_]+)\s*\(", text))


def test_library_exports_every_declared_symbol():
    from rgqa_b200 import _lib as L
    lib = L.load()
    declared = _header_symbols()
    assert declared, "no declarations parsed from the header"
    for name in declared:
        assert hasattr(lib, name), f"{name} is declared in include/rgqa_b200.h but not exported"
    assert set(L.SIGNATURES) == declared, set(L.SIGNATURES) ^ declared
    assert lib.rgqa_abi_version() == L.ABI_VERSION == 3


def test_no_cpu_path():
    """Without a CUDA device the product must refuse, not fall back."""
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from rgqa_b200 import _lib as L
    from rgqa_b200.engine import Engine, EngineConfig
    with pytest.raises(RuntimeError, match="CUDA device"):
        Engine(EngineConfig())
    lib = L.load()
    cfg = EngineConfig().to_c()
    h = C.c_void_p()
    assert lib.rgqa_create(C.byref(cfg), C.byref(h)) != 0
    assert b"no CUDA device" in lib.rgqa_last_error()
    from rgqa_b200 import ops
    with pytest.raises(RuntimeError, match="CUDA tensors required"):
        ops.layernorm(torch.zeros(2, 768), torch.ones(768), torch.zeros(768))


def test_product_never_imports_the_oracle():
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "rgqa_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                if re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M) or "oracle/" in src:
                    bad.append(os.path.join(dirpath, f))
    assert not bad, bad


def test_state_dict_schema_matches_reference_appendix_a(monkeypatch):
    monkeypatch.setattr(sys, "argv", ["x"])
    from rgqa_b200 import synthetic as syn
    from rgqa_b200.tasks.gqa_model import GQAModel, GQAModel_maha
    m = GQAModel(1842)
    sd = m.state_dict()
    assert len(sd) == 455 and sum(v.numel() for v in sd.values()) == 211952178
    assert "lxrt_encoder.model.bert.encoder.x_layers.4.visual_attention.att.key.weight" in sd
    assert sd["logit_fc.3.weight"].shape == (1842, 1536)
    ref = syn.lxmert_state_dict(variant="rich")
    assert m.load_state_dict(ref, strict=True).missing_keys == []
    for k, v in m.state_dict().items():
        assert torch.equal(v, ref[k]), k
    # DataParallel-style checkpoints: GQA.load strips '.module' (gqa_conf.py:378-380); non-strict load reports
    r = m.load_state_dict({k: v for k, v in list(ref.items())[:5]}, strict=False)
    assert len(r.missing_keys) == 450 and r.unexpected_keys == []
    with pytest.raises(RuntimeError, match="size mismatch"):
        m.load_state_dict({**ref, "logit_fc.3.bias": torch.zeros(7)}, strict=True)
    with pytest.raises(RuntimeError, match="Unexpected key"):
        m.load_state_dict({**ref, "bogus.weight": torch.zeros(1)}, strict=True)
    # fingerprint changes when a weight changes -> the packed engine is rebuilt
    fp = m.logit_fc.fingerprint()
    with torch.no_grad():
        next(m.logit_fc.parameters()).add_(1.0)
    assert m.logit_fc.fingerprint() != fp
    assert GQAModel_maha.returns_pooled and m.lxrt_encoder.dim == 768 and m.lxrt_encoder.max_seq_length == 30
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.zeros(1, 36, 2048), torch.zeros(1, 36, 4), ["what is this?"])


def test_param_flags_follow_reference_names(monkeypatch):
    from rgqa_b200 import param
    a = param.parse_args(["--test", "ClipEasy", "--load", "snap/x", "--batchSize", "32", "--llayers", "9", "--xlayers", "5",
                          "--rlayers", "5", "--tau", "0.3", "--noise", "0", "--mixup", "--unknownFlag", "1"])
    assert (a.test, a.load, a.batch_size, a.llayers, a.tau, a.precision) == ("ClipEasy", "snap/x", 32, 9, 0.3, "bf16")
    # the flags of the ensemble / ODIN / RP-finetune scripts (param.py:51-53, :60, :67, :88-91)
    b = param.parse_args(["--test", "GQAUQ_testdev", "--load", "snap/a,snap/b", "--ensemble_method", "multiply", "--sample_pair",
                          "--lr", "5e-5", "--epochs", "4", "--noise", "1e-4", "--temperature", "2", "--mixup_alpha", "0.3"])
    assert (b.ensemble_method, b.sample_pair, b.lr, b.epochs, b.noise, b.temperature) == ("multiply", True, 5e-5, 4, 1e-4, 2.0)
    assert b.load.split(",") == ["snap/a", "snap/b"] and b.optim == "bert" and b.dropout == 0.1 and not b.mce_loss
