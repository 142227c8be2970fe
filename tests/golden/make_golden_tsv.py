"""Generates tests/golden/obj_tiny.tsv (+ obj_tiny_expected.npz): a 4-image object-feature TSV in the bottom-up-attention
format the reference reads, and what the UNMODIFIED reference `load_obj_tsv` (src/utils.py:16-54) and
`GQATorchDataset.__getitem__`'s box normalisation (src/tasks/gqa_data.py:195-200, :240-250) make of it.
Run in the build container (needs /root/reference):  python tests/golden/make_golden_tsv.py"""
import base64, importlib.util, os, sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
rng = np.random.default_rng(20240)
NB, FD = 36, 64
rows = []
for i in range(4):
    w, h = int(rng.integers(300, 640)), int(rng.integers(200, 480))
    x1 = rng.uniform(0, w * 0.5, NB); y1 = rng.uniform(0, h * 0.5, NB)
    boxes = np.stack([x1, y1, x1 + rng.uniform(1, w * 0.5, NB), y1 + rng.uniform(1, h * 0.5, NB)], 1).astype(np.float32)
    feats = np.maximum(rng.normal(0.3, 1.0, (NB, FD)), 0).astype(np.float32)      # post-ReLU RoI features
    feats[0, :4] = [1.00390625, 1.01171875, 3.0e-39, 65504.0]                       # bf16 ties, a denormal, a large value
    rows.append({"img_id": f"n{1000 + i}", "img_h": h, "img_w": w,
                 "objects_id": rng.integers(0, 1600, NB).astype(np.int64), "objects_conf": rng.uniform(0, 1, NB).astype(np.float32),
                 "attrs_id": rng.integers(0, 400, NB).astype(np.int64), "attrs_conf": rng.uniform(0, 1, NB).astype(np.float32),
                 "num_boxes": NB, "boxes": boxes, "features": feats})
FIELDS = ["img_id", "img_h", "img_w", "objects_id", "objects_conf", "attrs_id", "attrs_conf", "num_boxes", "boxes", "features"]
with open(os.path.join(HERE, "obj_tiny.tsv"), "w") as f:
    for r in rows:
        f.write("\t".join(str(r[k]) if not isinstance(r[k], np.ndarray) else base64.b64encode(r[k].tobytes()).decode() for k in FIELDS) + "\n")

spec = importlib.util.spec_from_file_location("ref_utils", "/root/reference/src/utils.py")
ref = importlib.util.module_from_spec(spec); spec.loader.exec_module(ref)
data = ref.load_obj_tsv(os.path.join(HERE, "obj_tiny.tsv"))
out = {"img_ids": np.array([d["img_id"] for d in data])}
for k in ("img_h", "img_w", "num_boxes"):
    out[k] = np.array([d[k] for d in data])
for k in ("objects_id", "objects_conf", "attrs_id", "attrs_conf", "boxes", "features"):
    out[k] = np.stack([d[k] for d in data])
# the reference's per-item box handling, executed as written
norm, uni = [], []
for d in data:
    boxes = d["boxes"].copy()
    img_h, img_w = d["img_h"], d["img_w"]
    b = boxes.copy()
    b[:, (0, 2)] /= img_w
    b[:, (1, 3)] /= img_h
    norm.append(b)
    nb = np.zeros((boxes.shape[0], 7), dtype="float32")
    nb[:, 1] = boxes[:, 0]; nb[:, 0] = boxes[:, 1]; nb[:, 3] = boxes[:, 2]; nb[:, 2] = boxes[:, 3]
    nb[:, 4] = nb[:, 3] - nb[:, 1]; nb[:, 5] = nb[:, 2] - nb[:, 0]; nb[:, 6] = nb[:, 4] * nb[:, 5]
    uni.append(nb)
out["boxes_lxmert"] = np.stack(norm); out["boxes_uniter"] = np.stack(uni)
top2 = ref.load_obj_tsv(os.path.join(HERE, "obj_tiny.tsv"), topk=2)
out["topk2_ids"] = np.array([d["img_id"] for d in top2])
np.savez_compressed(os.path.join(HERE, "obj_tiny_expected.npz"), **out)
print("wrote", len(data), "images;", os.path.getsize(os.path.join(HERE, "obj_tiny.tsv")), "bytes of TSV")
