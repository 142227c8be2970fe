"""The host tokenizer must reproduce the reference BertTokenizer + convert_sents_to_features
(src/lxrt/tokenization.py, src/lxrt/entry.py:36-71) token for token; golden produced by the reference
itself over fixture questions plus adversarial strings (tests/golden/make_golden.py --only tokenizer)."""
import gzip
import json
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def gold():
    with gzip.open(os.path.join(GOLD, "tokenizer_golden.json.gz"), "rt") as f:
        return json.load(f)


def _tok():
    from rgqa_b200.lxrt.tokenization import BertTokenizer
    return BertTokenizer(os.path.join(GOLD, "tokenizer_vocab.txt"), do_lower_case=True)


def test_tokens_match_reference(gold):
    tok = _tok()
    for sent, want in zip(gold["sents"], gold["tokens"]):
        assert tok.tokenize(sent.strip()) == want, sent


@pytest.mark.parametrize("L", [20, 30])
def test_encode_batch_matches_convert_sents_to_features(gold, L):
    tok = _tok()
    ids, mask, seg = tok.encode_batch(gold["sents"], L)
    g = gold[str(L)]
    np.testing.assert_array_equal(ids, np.asarray(g["input_ids"]))
    np.testing.assert_array_equal(mask, np.asarray(g["input_mask"]))
    np.testing.assert_array_equal(seg, np.asarray(g["segment_ids"]))
    assert ids.dtype == np.int32


def test_roundtrip_and_errors(tmp_path):
    from rgqa_b200.lxrt.tokenization import BertTokenizer, SyntheticTokenizer
    tok = _tok()
    t = tok.tokenize("what color is the sky?")
    assert tok.convert_ids_to_tokens(tok.convert_tokens_to_ids(t)) == t
    with pytest.raises(KeyError):
        tok.convert_tokens_to_ids(["definitely-not-in-vocab"])
    with pytest.raises(ValueError):
        BertTokenizer(str(tmp_path / "missing.txt"))
    with pytest.raises(EnvironmentError):
        BertTokenizer.from_pretrained("bert-base-uncased")  # no network, no local vocab
    s = SyntheticTokenizer()
    ids, mask, _ = s.encode_batch(["Is the cat on the mat?", ""], 20)
    assert ids[0, 0] == 101 and ids[0, mask[0].sum() - 1] == 102 and ids[1, 1] == 102
    assert (ids[0, 1:mask[0].sum() - 1] >= 1000).all() and ids.max() < 30522


def test_tokenizer_fallback_is_loud_and_refused_with_checkpoints(monkeypatch):
    """the hash-based SyntheticTokenizer must never sit silently in front of trained embeddings.  Without a
    vocabulary: a warning for checkpoint-free models, an error as soon as a checkpoint is involved; `--vocab synthetic` is
    the explicit opt-in; a real vocab file resolves to the WordPiece tokenizer."""
    import types
    import warnings
    import pytest
    from rgqa_b200.lxrt import entry
    from rgqa_b200.lxrt.tokenization import BertTokenizer, SyntheticTokenizer
    monkeypatch.delenv("RGQA_BERT_VOCAB_DIR", raising=False)
    ns = lambda **kw: types.SimpleNamespace(**{"vocab": None, "load": None, "load_lxmert": None, "load_lxmert_qa": None, **kw})
    with pytest.warns(UserWarning, match="SyntheticTokenizer"):
        tok = entry.make_tokenizer(ns())
    assert isinstance(tok, SyntheticTokenizer) and tok.is_fallback
    with pytest.raises(EnvironmentError, match="vocabulary"):
        entry.require_real_vocab(tok, "GQA.load")
    with pytest.raises(EnvironmentError, match="checkpoint is being loaded"):
        entry.make_tokenizer(ns(load="snap/gqa/x/BEST"))
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        explicit = entry.make_tokenizer(ns(vocab="synthetic", load="snap/gqa/x/BEST"))
    assert isinstance(explicit, SyntheticTokenizer) and not getattr(explicit, "is_fallback", False)
    entry.require_real_vocab(explicit, "GQA.load")
    vocab = os.path.join(os.path.dirname(__file__), "golden", "tokenizer_vocab.txt")
    real = entry.make_tokenizer(ns(vocab=vocab, load="snap/gqa/x/BEST"))
    assert isinstance(real, BertTokenizer)
