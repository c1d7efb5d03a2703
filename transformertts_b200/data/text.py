"""Phoneme-symbol table and tokenizer of the reference (data/text/symbols.py, data/text/tokenizer.py:9-49), needed to turn
the phonemized metadata files of the training-data directory into the token ids the model was trained on.  The symbol
inventory is DATA (the ids of a published checkpoint depend on it character for character); the espeak phonemizer itself
(text -> phonemes) is outside the text->mel hot path and is not reproduced.

Checked against the reference classes in tests/test_reference_shim.py."""
from __future__ import annotations

from typing import List

# IPA inventory, grouped as in the reference's symbol table
VOWELS = 'iyɨʉɯuɪʏʊeøɘəɵɤoɛœɜɞʌɔæɐaɶɑɒᵻ'
NON_PULMONIC = 'ʘɓǀɗǃʄǂɠǁʛ'
PULMONIC = 'pbtdʈɖcɟkɡqɢʔɴŋɲɳnɱmʙrʀⱱɾɽɸβfvθðszʃʒʂʐçʝxɣχʁħʕhɦɬɮʋɹɻjɰlɭʎʟ'
SUPRASEGMENTALS = 'ˈˌːˑ'
OTHER = 'ʍwɥʜʢʡɕʑɺɧ'
DIACRITICS = 'ɚ˞ɫ'
PUNCTUATION = '!,-.:;? \'()'

ALL_PHONEMES: List[str] = sorted(list(sorted(list(VOWELS + NON_PULMONIC + PULMONIC + SUPRASEGMENTALS + OTHER + DIACRITICS))) + list(PUNCTUATION))


class Tokenizer:
    """ids: 0 = padding, 1..len(alphabet) = symbols in sorted order, then (optionally) start, end and breathing tokens."""

    def __init__(self, start_token: str = '>', end_token: str = '<', pad_token: str = '/', add_start_end: bool = True, alphabet=None,
                 model_breathing: bool = True):
        self.alphabet = sorted(set(alphabet)) if alphabet else list(ALL_PHONEMES)
        self.idx_to_token = {0: pad_token}
        self.idx_to_token.update({i: s for i, s in enumerate(self.alphabet, start=1)})
        self.token_to_idx = {s: [i] for i, s in self.idx_to_token.items()}
        self.vocab_size = len(self.alphabet) + 1
        self.add_start_end = add_start_end
        if add_start_end:
            self.start_token_index, self.end_token_index = self.vocab_size, self.vocab_size + 1
            self.idx_to_token[self.start_token_index], self.idx_to_token[self.end_token_index] = start_token, end_token
            self.vocab_size += 2
        self.model_breathing = model_breathing
        if model_breathing:   # every blank is followed by a breathing token, and one opens the sentence
            self.breathing_token, self.breathing_token_index = '@', self.vocab_size
            self.token_to_idx[' '] = self.token_to_idx[' '] + [self.breathing_token_index]
            self.idx_to_token[self.breathing_token_index] = self.breathing_token
            self.token_to_idx[self.breathing_token] = [self.breathing_token_index]
            self.vocab_size += 1

    def __call__(self, sentence: str) -> List[int]:
        seq = [i for c in sentence for i in self.token_to_idx[c]]   # unknown characters raise KeyError, as in the reference
        if self.model_breathing:
            seq = [self.breathing_token_index] + seq
        if self.add_start_end:
            seq = [self.start_token_index] + seq + [self.end_token_index]
        return seq

    def decode(self, sequence) -> str:
        return ''.join(self.idx_to_token[int(t)] for t in sequence)


class TextToTokens:
    """data/text/__init__.py:7-23 with the phonemizer injected (a callable text -> phoneme string; None = input is phonemes)."""

    def __init__(self, tokenizer: Tokenizer, phonemizer=None):
        self.tokenizer, self.phonemizer = tokenizer, phonemizer

    def __call__(self, text):
        return self.tokenizer(self.phonemizer(text) if self.phonemizer is not None else text)

    @classmethod
    def default(cls, language: str = 'en-us', add_start_end: bool = False, with_stress: bool = True, model_breathing: bool = False,
                phonemizer=None):
        return cls(Tokenizer(add_start_end=add_start_end, model_breathing=model_breathing), phonemizer)
