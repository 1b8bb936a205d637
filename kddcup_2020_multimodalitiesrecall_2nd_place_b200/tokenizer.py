"""WordPiece tokenisation of queries and class-label phrases.

Behaviour of the reference tokenizers (imagebert_zk/tokenization.py:161-359, identical algorithm in
lxmert/src/lxrt/tokenization.py:72-348): text -> clean (drop NUL / U+FFFD / control characters, every whitespace
becomes a space) -> CJK characters isolated by spaces -> whitespace split -> [lower-case + strip combining marks] ->
split on punctuation -> greedy longest-match-first WordPiece with the "##" continuation prefix, words longer than 200
(TF trees) / 100 (lxmert) characters or without a full cover become [UNK].  Written from that specification; pinned against the reference's own
classes by tests/golden/tokenizer_kat.json (tools/make_golden.py).
"""
from __future__ import annotations

import unicodedata
from typing import Dict, Iterable, List


def load_vocab(vocab_file: str) -> Dict[str, int]:
    """One token per line, id = line number."""
    vocab: Dict[str, int] = {}
    with open(vocab_file, encoding="utf-8") as f:
        for i, line in enumerate(f):
            tok = line.rstrip("\n").strip()
            if tok not in vocab:           # first occurrence wins (an OrderedDict assignment would keep the last; the
                vocab[tok] = i             # shipped vocab.txt has no duplicates)
    return vocab


def _is_whitespace(ch: str) -> bool:
    return ch in (" ", "\t", "\n", "\r") or unicodedata.category(ch) == "Zs"


def _is_control(ch: str) -> bool:
    if ch in ("\t", "\n", "\r"):
        return False
    return unicodedata.category(ch) in ("Cc", "Cf")


def _is_punctuation(ch: str) -> bool:
    cp = ord(ch)
    if 33 <= cp <= 47 or 58 <= cp <= 64 or 91 <= cp <= 96 or 123 <= cp <= 126:   # all non-alphanumeric ASCII
        return True
    return unicodedata.category(ch).startswith("P")


def _is_cjk(cp: int) -> bool:
    return (0x4E00 <= cp <= 0x9FFF or 0x3400 <= cp <= 0x4DBF or 0x20000 <= cp <= 0x2A6DF or 0x2A700 <= cp <= 0x2B73F
            or 0x2B740 <= cp <= 0x2B81F or 0x2B820 <= cp <= 0x2CEAF or 0xF900 <= cp <= 0xFAFF or 0x2F800 <= cp <= 0x2FA1F)


class BasicTokenizer:
    def __init__(self, do_lower_case: bool = True):
        self.do_lower_case = do_lower_case

    def tokenize(self, text: str) -> List[str]:
        if isinstance(text, bytes):
            text = text.decode("utf-8", "ignore")
        cleaned = []
        for ch in text:
            cp = ord(ch)
            if cp == 0 or cp == 0xFFFD or _is_control(ch):
                continue
            if _is_whitespace(ch):
                cleaned.append(" ")
            elif _is_cjk(cp):
                cleaned.append(" " + ch + " ")
            else:
                cleaned.append(ch)
        out: List[str] = []
        for word in "".join(cleaned).split():
            if self.do_lower_case:
                word = "".join(c for c in unicodedata.normalize("NFD", word.lower()) if unicodedata.category(c) != "Mn")
            piece = []
            for ch in word:
                if _is_punctuation(ch):
                    if piece:
                        out.append("".join(piece))
                        piece = []
                    out.append(ch)
                else:
                    piece.append(ch)
            if piece:
                out.append("".join(piece))
        return " ".join(out).split()


class WordpieceTokenizer:
    def __init__(self, vocab: Dict[str, int], unk_token: str = "[UNK]", max_input_chars_per_word: int = 200):
        self.vocab, self.unk_token, self.max_chars = vocab, unk_token, max_input_chars_per_word

    def tokenize(self, text: str) -> List[str]:
        out: List[str] = []
        for word in text.split():
            if len(word) > self.max_chars:
                out.append(self.unk_token)
                continue
            pieces, start, ok = [], 0, True
            while start < len(word):
                end = len(word)
                found = None
                while start < end:
                    sub = word[start:end] if start == 0 else "##" + word[start:end]
                    if sub in self.vocab:
                        found = sub
                        break
                    end -= 1
                if found is None:
                    ok = False
                    break
                pieces.append(found)
                start = end
            out.extend(pieces if ok else [self.unk_token])
        return out


class FullTokenizer:
    """tokenization.FullTokenizer(vocab_file, do_lower_case) of the reference (tokenization.py:161-183)."""

    def __init__(self, vocab_file: str = None, do_lower_case: bool = True, vocab: Dict[str, int] = None,
                 max_input_chars_per_word: int = 200):
        """max_input_chars_per_word: 200 in the two TF trees (tokenization.py:303), 100 in lxmert's
        (lxrt/tokenization.py:294)."""
        self.vocab = vocab if vocab is not None else load_vocab(vocab_file)
        self.inv_vocab = {v: k for k, v in self.vocab.items()}
        self.basic_tokenizer = BasicTokenizer(do_lower_case=do_lower_case)
        self.wordpiece_tokenizer = WordpieceTokenizer(vocab=self.vocab, max_input_chars_per_word=max_input_chars_per_word)

    def tokenize(self, text: str) -> List[str]:
        out: List[str] = []
        for token in self.basic_tokenizer.tokenize(text):
            out.extend(self.wordpiece_tokenizer.tokenize(token))
        return out

    def convert_tokens_to_ids(self, tokens: Iterable[str]) -> List[int]:
        return [self.vocab[t] for t in tokens]       # KeyError on an out-of-vocabulary token, as in the reference

    def convert_ids_to_tokens(self, ids: Iterable[int]) -> List[str]:
        return [self.inv_vocab[i] for i in ids]
