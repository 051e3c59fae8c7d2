"""Stand-in for the T5 sentencepiece tokenizer (TEST INFRASTRUCTURE).  There is no `spiece.model` offline,
so the reference's host plumbing (Collator, TestDataset, the candidate Trie) is driven with this small deterministic
tokenizer exposing exactly the calls the reference makes:
    batch_encode_plus(texts, padding="longest", truncation=True, max_length=512)   processor/Collator.py:12-21
    convert_ids_to_tokens(ids)                                                      processor/Collator.py:17
    encode(text)                                                                    runner/DistributedRunner.py:346
    batch_decode(ids, skip_special_tokens=True)                                     runner/DistributedRunner.py:376-381
    add_tokens(list), len(tokenizer)                                                main.py:188-193
Pieces mimic sentencepiece: the first piece of every whitespace-separated word carries the "▁" marker; a word is cut at
letter / digit / other boundaries, digit runs in chunks of three ("item_1001" -> "▁item", "_", "100", "1").  Ids are
assigned on first sight (deterministic for a fixed call order); pad = 0, eos = 1, unk = 2, like T5."""
import re


class StandInTokenizer:
    pad_token_id, eos_token_id, unk_token_id = 0, 1, 2

    def __init__(self, vocab_size=32100, first_id=3):
        self.vocab = {"<pad>": 0, "</s>": 1, "<unk>": 2}
        self.inv = {0: "<pad>", 1: "</s>", 2: "<unk>"}
        self.vocab_size = vocab_size
        self.next_id = first_id
        self.added = []

    def __len__(self):
        return self.vocab_size + len(self.added)

    def add_tokens(self, tokens):
        n = 0
        for t in tokens:
            if t not in self.vocab:
                tid = self.vocab_size + len(self.added)
                self.vocab[t] = tid
                self.inv[tid] = t
                self.added.append(t)
                n += 1
        return n

    def _pieces(self, text):
        out = []
        for word in text.split():
            # added tokens (e.g. <CI12>) are kept whole, as HF does for added vocabulary
            parts = re.findall(r"<[^<>\s]+>|[A-Za-z]+|[0-9]{1,3}|[^A-Za-z0-9\s]", word)
            for i, p in enumerate(parts):
                out.append(("▁" + p) if i == 0 else p)
        return out

    def _id(self, piece):
        bare = piece[1:] if piece.startswith("▁") else piece
        if bare in self.vocab and bare.startswith("<"):      # added token: one id with or without the word marker
            return self.vocab[bare]
        if piece not in self.vocab:
            assert self.next_id < self.vocab_size, "stand-in vocabulary exhausted"
            self.vocab[piece] = self.next_id
            self.inv[self.next_id] = piece
            self.next_id += 1
        return self.vocab[piece]

    def tokenize(self, text):
        return self._pieces(text)

    def encode(self, text, add_special_tokens=True):
        ids = [self._id(p) for p in self._pieces(text)]
        return ids + [1] if add_special_tokens else ids

    def batch_encode_plus(self, texts, padding="longest", truncation=True, max_length=512, **_):
        rows = []
        for t in texts:
            ids = self.encode(t)
            if truncation and len(ids) > max_length:
                ids = ids[: max_length - 1] + [1]
            rows.append(ids)
        L = max(len(r) for r in rows) if padding else None
        input_ids = [r + [0] * (L - len(r)) for r in rows] if padding else rows
        mask = [[1] * len(r) + [0] * (L - len(r)) for r in rows] if padding else [[1] * len(r) for r in rows]
        return {"input_ids": input_ids, "attention_mask": mask}

    def __call__(self, texts, **kw):
        return self.batch_encode_plus(texts, **kw)

    def convert_ids_to_tokens(self, ids):
        return [self.inv.get(int(i), "<unk>") for i in ids]

    def decode(self, ids, skip_special_tokens=True):
        toks = []
        for i in ids:
            i = int(i)
            if skip_special_tokens and i in (0, 1):
                continue
            toks.append(self.inv.get(i, "<unk>"))
        return "".join(toks).replace("▁", " ").strip()

    def batch_decode(self, batch, skip_special_tokens=True):
        return [self.decode(row.tolist() if hasattr(row, "tolist") else row, skip_special_tokens) for row in batch]
