# make_prg/from_msa/__init__.py:2-3
NESTING_LVL = 5
MIN_MATCH_LEN = 7
