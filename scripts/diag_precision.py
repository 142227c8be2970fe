"""Print bf16-path error statistics against the reference golden logits (GPU box)."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rgqa_b200 import synthetic as syn
from rgqa_b200.engine import Engine, EngineConfig
GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
for tag in ["lxmert_init_L20_B4", "lxmert_rich_L30_B4", "lxmert_rich_L20_B6_head02", "uniter_init_L20_B4", "uniter_rich_L20_B4"]:
    z = np.load(os.path.join(GOLD, tag + ".npz")); meta = json.loads(str(z["meta"]))
    arch = tag.split("_")[0]
    cfg = EngineConfig(arch=arch, precision="bf16")
    if arch == "uniter":
        cfg.l_layers, cfg.x_layers, cfg.r_layers, cfg.vocab_size, cfg.visual_pos_dim = 12, 0, 0, 28996, 7
    sd = (syn.lxmert_state_dict if arch == "lxmert" else syn.uniter_state_dict)(variant=meta["variant"], head_sigma=meta.get("head_sigma"))
    e = Engine(cfg); e.load_weights(sd.items())
    inp = syn.make_inputs(meta["B"], meta["L"], seed=meta["inputs_seed"], vocab=cfg.vocab_size)
    d = e.device
    pooled, logits = e.forward(inp["feat"].to(d), inp["pos7" if arch == "uniter" else "pos"].to(d), inp["ids"].to(d), None, inp["mask"].to(d))
    dl = logits.cpu().numpy() - z["logits"]; dp = pooled.cpu().numpy() - z["pooled"]
    print(f"{tag:28s} logits: max {np.abs(dl).max():.4f} rms {np.sqrt((dl**2).mean()):.4f} (ref std {z['logits'].std():.3f})  "
          f"pooled: max {np.abs(dp).max():.4f} rms {np.sqrt((dp**2).mean()):.5f}  argmax agree {(logits.cpu().numpy().argmax(1)==z['logits'].argmax(1)).mean():.2f}")
    e.close()
