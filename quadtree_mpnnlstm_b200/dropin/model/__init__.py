"""Drop-in ``model`` package: put ``quadtree_mpnnlstm_b200/dropin`` (and the repo root) on PYTHONPATH and the
reference's ``ice_exp.py`` / ``ice_inf.py`` / notebook imports (``from model.seq2seq import Seq2Seq`` ...)
resolve to the B200 implementation.  See INTEGRATION.md."""
