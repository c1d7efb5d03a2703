"""Import stub: the espeak phonemizer is outside the text->mel path (token ids are the model input in every test)."""
