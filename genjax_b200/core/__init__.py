"""Host-side core datatypes: PRNG keys, choice maps, selections."""
