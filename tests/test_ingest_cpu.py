"""Feature ingest on the CPU: the TSV reader against what the UNMODIFIED reference `load_obj_tsv`
produced for tests/golden/obj_tiny.tsv (fixture made by tests/golden/make_golden_tsv.py), the one-time conversion to
the memory-mapped store, and the sweep loader's batches."""
import base64
import os

import numpy as np
import pytest
import torch

from rgqa_b200 import ingest
from rgqa_b200.lxrt.tokenization import BertTokenizer

GOLD = os.path.join(os.path.dirname(__file__), "golden")
TSV = os.path.join(GOLD, "obj_tiny.tsv")


@pytest.fixture(scope="module")
def expected():
    return np.load(os.path.join(GOLD, "obj_tiny_expected.npz"))


def test_load_obj_tsv_matches_reference_loader(expected):
    data = ingest.load_obj_tsv(TSV)
    assert [d["img_id"] for d in data] == list(expected["img_ids"])
    for k in ("img_h", "img_w", "num_boxes"):
        assert [d[k] for d in data] == list(expected[k]) and all(isinstance(d[k], int) for d in data)
    for k in ("objects_id", "objects_conf", "attrs_id", "attrs_conf", "boxes", "features"):
        got = np.stack([d[k] for d in data])
        assert got.dtype == expected[k].dtype and np.array_equal(got, expected[k])      # bit-exact decode
        assert not data[0][k].flags.writeable                                           # src/utils.py:48
    assert [d["img_id"] for d in ingest.load_obj_tsv(TSV, topk=2)] == list(expected["topk2_ids"])
    assert len(ingest.load_obj_tsv(TSV, topk=-1)) == 4


def test_box_forms_match_reference_getitem(expected):
    data = ingest.load_obj_tsv(TSV)
    for i, d in enumerate(data):
        assert np.array_equal(ingest.normalize_boxes(d["boxes"], d["img_w"], d["img_h"]), expected["boxes_lxmert"][i])
        assert np.array_equal(ingest.uniter_boxes(d["boxes"]), expected["boxes_uniter"][i])
    bad = np.array([[0, 0, 700, 10]], dtype=np.float32)
    with pytest.raises(AssertionError):                       # the reference asserts boxes inside the image (:199-200)
        ingest.normalize_boxes(bad, 640, 480)


def test_bf16_conversion_is_round_to_nearest_even():
    rng = np.random.default_rng(3)
    a = np.concatenate([rng.normal(0, 1, 50000), rng.normal(0, 1e-30, 1000), rng.normal(0, 1e30, 1000),
                        [0.0, -0.0, 1.00390625, 1.01171875, np.inf, -np.inf, 3.0e-39, 65504.0]]).astype(np.float32)
    bits = ingest.f32_to_bf16_bits(a)
    ref = torch.from_numpy(a).to(torch.bfloat16).view(torch.int16).numpy().view(np.uint16)
    assert np.array_equal(bits, ref)
    nan = ingest.f32_to_bf16_bits(np.array([np.nan], dtype=np.float32))
    assert (nan[0] & 0x7F80) == 0x7F80 and (nan[0] & 0x7F) != 0


@pytest.mark.parametrize("dtype", ["bf16", "f32"])
def test_convert_and_store_round_trip(tmp_path, expected, dtype):
    store = ingest.convert_obj_tsv(TSV, str(tmp_path / dtype), dtype=dtype)
    assert len(store) == 4 and "n1002" in store and "zzz" not in store
    again = ingest.FeatureStore(str(tmp_path / dtype))       # re-open from disk
    feat = torch.empty((3, 36, 64), dtype=again.torch_dtype)
    box = torch.empty((3, 36, 4), dtype=torch.float32)
    again.gather(["n1003", "n1000", "n1003"], feat, box)
    want = torch.from_numpy(expected["features"][[3, 0, 3]].copy())
    want = want.to(torch.bfloat16) if dtype == "bf16" else want
    assert torch.equal(feat, want)
    assert np.array_equal(box.numpy(), expected["boxes_lxmert"][[3, 0, 3]])
    uni = ingest.convert_obj_tsv(TSV, str(tmp_path / (dtype + "_u")), dtype=dtype, backbone="uniter", topk=2)
    assert len(uni) == 2 and np.array_equal(np.asarray(uni.boxes), expected["boxes_uniter"][:2])


def test_convert_rejects_ragged_box_counts(tmp_path):
    rows = ingest.load_obj_tsv(TSV, topk=2)
    fields = ingest.FIELDNAMES
    def line(r, nb):
        vals = dict(r, num_boxes=nb)
        for k in ("objects_id", "objects_conf", "attrs_id", "attrs_conf", "boxes", "features"):
            vals[k] = base64.b64encode(np.ascontiguousarray(r[k][:nb]).tobytes()).decode()
        return "\t".join(str(vals[k]) for k in fields) + "\n"
    p = tmp_path / "ragged.tsv"
    p.write_text(line(rows[0], 36) + line(rows[1], 20))
    with pytest.raises(ValueError):
        ingest.convert_obj_tsv(str(p), str(tmp_path / "out"))


def test_sweep_loader_batches_in_dataset_order(tmp_path):
    store = ingest.convert_obj_tsv(TSV, str(tmp_path / "s"), dtype="bf16")
    tok = BertTokenizer(os.path.join(GOLD, "tokenizer_vocab.txt"), do_lower_case=True)
    sents = ["is the cat on the mat?", "what color is the bus", "", "are there any zebras to the left of the tree?",
             "who is holding the umbrella", "is it raining", "how many dogs"]
    imgs = ["n1001", "n1003", "n9999", "n1000", "n1001", "n1002", "n1003"]          # n9999 has no features: dropped
    data = [{"question_id": f"q{i}", "img_id": im, "sent": s} for i, (im, s) in enumerate(zip(imgs, sents))]
    loader = ingest.SweepLoader(store, data, tok, max_seq_length=20, batch_size=4)
    assert len(loader) == 2
    batches = [(q, f.clone(), b.clone(), i.clone(), m.clone()) for q, f, b, i, m in loader]
    assert [q for b in batches for q in b[0]] == ["q0", "q1", "q3", "q4", "q5", "q6"]
    kept = [d for d in data if d["img_id"] in store]
    ids, mask, _ = tok.encode_batch([d["sent"] for d in kept], 20)
    assert np.array_equal(torch.cat([b[3] for b in batches]).numpy(), ids)
    assert np.array_equal(torch.cat([b[4] for b in batches]).numpy(), mask)
    feats = torch.cat([b[1] for b in batches])
    rows = [store.imgid2row[d["img_id"]] for d in kept]
    assert feats.dtype == torch.bfloat16
    assert np.array_equal(feats.view(torch.int16).numpy().view(np.uint16), np.asarray(store.features)[rows])
    # a rank's shard
    shard = ingest.SweepLoader(store, data, tok, max_seq_length=20, batch_size=4, lo=2, hi=5)
    assert [q for qs, *_ in shard for q in qs] == ["q3", "q4", "q5"]
