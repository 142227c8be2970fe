"""BERT basic + WordPiece tokenisation for the question side of the hot path (host code).

Same public surface as the reference's ``src/lxrt/tokenization.py`` (``BertTokenizer`` :72 with
``tokenize`` / ``convert_tokens_to_ids`` / ``convert_ids_to_tokens`` / ``from_pretrained``), same
output on every input (pinned by tests/golden/tokenizer_golden.json.gz, produced by the reference
tokenizer), different construction: one pass over the text driven by a memoised per-character class
table, and a per-word WordPiece cache — GQA questions reuse a few thousand words, so a sweep costs a
dictionary lookup per word instead of the reference's per-character Python loops
(≈ 7 k questions/s measured in SURVEY.md §7, which would otherwise cap the GPU path).

``encode_batch`` is the batched form of ``convert_sents_to_features`` (src/lxrt/entry.py:36-71):
``[CLS] + tokens[:L-2] + [SEP]``, zero padded to L, returning int32 arrays ready for one H2D copy.
"""
from __future__ import annotations

import collections
import os
import unicodedata

import numpy as np

VOCAB_NAME = "vocab.txt"
_DROP, _SPACE, _PUNCT, _CJK, _PLAIN = range(5)
_CJK_RANGES = ((0x4E00, 0x9FFF), (0x3400, 0x4DBF), (0x20000, 0x2A6DF), (0x2A700, 0x2B73F), (0x2B740, 0x2B81F),
               (0x2B820, 0x2CEAF), (0xF900, 0xFAFF), (0x2F800, 0x2FA1F))


def load_vocab(vocab_file):
    """token -> index, one token per line, stripped (duplicates: the last index wins, as in the reference)."""
    vocab = collections.OrderedDict()
    with open(vocab_file, "r", encoding="utf-8") as reader:
        for index, line in enumerate(reader):
            vocab[line.strip()] = index
    return vocab


def whitespace_tokenize(text):
    text = text.strip()
    return text.split() if text else []


class _CharClasses(dict):
    """char -> class, filled lazily (dict.__missing__), so each distinct character is classified once."""

    def __missing__(self, ch):
        cp = ord(ch)
        cat = unicodedata.category(ch)
        if ch in "\t\n\r" or ch == " " or cat == "Zs":
            c = _SPACE
        elif cp == 0 or cp == 0xFFFD or cat.startswith("C"):
            c = _DROP
        elif any(lo <= cp <= hi for lo, hi in _CJK_RANGES):
            c = _CJK
        elif (33 <= cp <= 47) or (58 <= cp <= 64) or (91 <= cp <= 96) or (123 <= cp <= 126) or cat.startswith("P"):
            c = _PUNCT
        else:
            c = _PLAIN
        self[ch] = c
        return c


_CLASSES = _CharClasses()


def _strip_accents(text):
    return "".join(ch for ch in unicodedata.normalize("NFD", text) if unicodedata.category(ch) != "Mn")


class BasicTokenizer(object):
    """Clean-up, CJK isolation, lower-casing + accent stripping, punctuation splitting."""

    def __init__(self, do_lower_case=True, never_split=("[UNK]", "[SEP]", "[PAD]", "[CLS]", "[MASK]")):
        self.do_lower_case = do_lower_case
        self.never_split = never_split

    def _split_punct(self, token):
        if token in self.never_split:
            return [token]
        out, cur = [], []
        for ch in token:
            if _CLASSES[ch] == _PUNCT:
                if cur:
                    out.append("".join(cur))
                    cur = []
                out.append(ch)
            else:
                cur.append(ch)
        if cur:
            out.append("".join(cur))
        return out

    def tokenize(self, text):
        # pass 1: drop control chars, normalise whitespace, put spaces around CJK characters
        buf = []
        for ch in text:
            c = _CLASSES[ch]
            if c == _DROP:
                continue
            if c == _SPACE:
                buf.append(" ")
            elif c == _CJK:
                buf.append(" " + ch + " ")
            else:
                buf.append(ch)
        pieces = []
        for token in "".join(buf).split():
            if self.do_lower_case and token not in self.never_split:
                token = _strip_accents(token.lower())
            pieces.extend(self._split_punct(token))
        # accent stripping can leave empty or space-bearing pieces; re-split like the reference does
        return " ".join(pieces).split()


class WordpieceTokenizer(object):
    """Greedy longest-match-first sub-word split with a per-word cache."""

    def __init__(self, vocab, unk_token="[UNK]", max_input_chars_per_word=100):
        self.vocab = vocab
        self.unk_token = unk_token
        self.max_input_chars_per_word = max_input_chars_per_word
        self._cache = {}

    def _split_word(self, word):
        hit = self._cache.get(word)
        if hit is not None:
            return hit
        n = len(word)
        if n > self.max_input_chars_per_word:
            out = (self.unk_token,)
        else:
            out, start = [], 0
            while start < n:
                end = n
                piece = None
                while start < end:
                    cand = word[start:end] if start == 0 else "##" + word[start:end]
                    if cand in self.vocab:
                        piece = cand
                        break
                    end -= 1
                if piece is None:
                    out = [self.unk_token]
                    break
                out.append(piece)
                start = end
            out = tuple(out)
        self._cache[word] = out
        return out

    def tokenize(self, text):
        out = []
        for word in whitespace_tokenize(text):
            out.extend(self._split_word(word))
        return out


class BertTokenizer(object):
    """End-to-end tokenisation: punctuation splitting + wordpiece (reference: tokenization.py:72-172)."""

    def __init__(self, vocab_file, do_lower_case=True, max_len=None, do_basic_tokenize=True,
                 never_split=("[UNK]", "[SEP]", "[PAD]", "[CLS]", "[MASK]")):
        if not os.path.isfile(vocab_file):
            raise ValueError(
                "Can't find a vocabulary file at path '{}'. To load the vocabulary from a Google pretrained "
                "model use `tokenizer = BertTokenizer.from_pretrained(PRETRAINED_MODEL_NAME)`".format(vocab_file))
        self.vocab = load_vocab(vocab_file)
        self.ids_to_tokens = collections.OrderedDict((i, t) for t, i in self.vocab.items())
        self.do_basic_tokenize = do_basic_tokenize
        if do_basic_tokenize:
            self.basic_tokenizer = BasicTokenizer(do_lower_case=do_lower_case, never_split=never_split)
        self.wordpiece_tokenizer = WordpieceTokenizer(vocab=self.vocab)
        self.max_len = max_len if max_len is not None else int(1e12)

    def tokenize(self, text):
        if not self.do_basic_tokenize:
            return self.wordpiece_tokenizer.tokenize(text)
        out = []
        for token in self.basic_tokenizer.tokenize(text):
            out.extend(self.wordpiece_tokenizer._split_word(token))
        return out

    def convert_tokens_to_ids(self, tokens):
        return [self.vocab[t] for t in tokens]  # KeyError on an unknown token, like the reference

    def convert_ids_to_tokens(self, ids):
        return [self.ids_to_tokens[i] for i in ids]

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, cache_dir=None, *inputs, **kwargs):
        """Accepts a vocab file or a directory holding ``vocab.txt``.  There is no network here, so the
        reference's S3 model names (tokenization.py:27-35) resolve only through ``$RGQA_BERT_VOCAB_DIR/<name>/vocab.txt``."""
        path = pretrained_model_name_or_path
        if not os.path.exists(path):
            root = os.environ.get("RGQA_BERT_VOCAB_DIR")
            if root and os.path.exists(os.path.join(root, path)):
                path = os.path.join(root, path)
        if os.path.isdir(path):
            path = os.path.join(path, VOCAB_NAME)
        if not os.path.isfile(path):
            raise EnvironmentError(
                "vocabulary '{}' not found: pass a vocab.txt path / directory or set RGQA_BERT_VOCAB_DIR "
                "(no download is attempted)".format(pretrained_model_name_or_path))
        kwargs['max_len'] = min(kwargs.get('max_len', int(1e12)), 512)
        return cls(path, *inputs, **kwargs)

    # ---- batched convert_sents_to_features (src/lxrt/entry.py:36-71) ----
    def encode_batch(self, sents, max_seq_length):
        n = len(sents)
        ids = np.zeros((n, max_seq_length), dtype=np.int32)
        mask = np.zeros((n, max_seq_length), dtype=np.int32)
        seg = np.zeros((n, max_seq_length), dtype=np.int32)
        cls_id, sep_id = self.vocab["[CLS]"], self.vocab["[SEP]"]
        vocab = self.vocab
        for r, sent in enumerate(sents):
            toks = self.tokenize(sent.strip())[:max_seq_length - 2]
            k = len(toks) + 2
            ids[r, 0] = cls_id
            if toks:
                ids[r, 1:k - 1] = [vocab[t] for t in toks]
            ids[r, k - 1] = sep_id
            mask[r, :k] = 1
        return ids, mask, seg


class SyntheticTokenizer(object):
    """Stand-in when no vocabulary file exists offline (random-init / synthetic benchmarks): hashes each
    basic token into [1000, vocab_size) deterministically.  Same interface as BertTokenizer."""

    def __init__(self, vocab_size=30522, do_lower_case=True):
        self.vocab_size = vocab_size
        self.basic_tokenizer = BasicTokenizer(do_lower_case=do_lower_case)
        self.vocab = {"[PAD]": 0, "[UNK]": 100, "[CLS]": 101, "[SEP]": 102, "[MASK]": 103}

    def tokenize(self, text):
        return self.basic_tokenizer.tokenize(text)

    def _id(self, tok):
        if tok in self.vocab:
            return self.vocab[tok]
        h = 2166136261
        for b in tok.encode("utf-8"):
            h = ((h ^ b) * 16777619) & 0xFFFFFFFF
        return 1000 + h % (self.vocab_size - 1000)

    def convert_tokens_to_ids(self, tokens):
        return [self._id(t) for t in tokens]

    def encode_batch(self, sents, max_seq_length):
        n = len(sents)
        ids = np.zeros((n, max_seq_length), dtype=np.int32)
        mask = np.zeros((n, max_seq_length), dtype=np.int32)
        seg = np.zeros((n, max_seq_length), dtype=np.int32)
        for r, sent in enumerate(sents):
            toks = self.tokenize(sent.strip())[:max_seq_length - 2]
            row = [101] + [self._id(t) for t in toks] + [102]
            ids[r, :len(row)] = row
            mask[r, :len(row)] = 1
        return ids, mask, seg
