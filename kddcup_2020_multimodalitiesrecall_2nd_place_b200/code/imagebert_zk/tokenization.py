"""imagebert_zk/tokenization.py of the reference (also imagebert_lds/src/tokenization.py and
lxmert/src/lxrt/tokenization.py:72-348): same class and function names over ..tokenizer."""
from ...tokenizer import BasicTokenizer, FullTokenizer, WordpieceTokenizer, load_vocab  # noqa: F401


def convert_tokens_to_ids(vocab, tokens):
    return [vocab[t] for t in tokens]


def convert_ids_to_tokens(inv_vocab, ids):
    return [inv_vocab[i] for i in ids]


def whitespace_tokenize(text):
    text = text.strip()
    return text.split() if text else []
