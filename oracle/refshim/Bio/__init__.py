"""Minimal stand-in for the Biopython surface that make_prg's from_msa path touches.

TEST INFRASTRUCTURE ONLY. It exists so that the *unmodified* reference source under
/root/reference can be imported in a container without Biopython, to validate the
oracle restatement and to generate the golden vectors in tests/golden/ (see
oracle/run_reference.py). Nothing in make_prg_b200/ imports it.
"""
