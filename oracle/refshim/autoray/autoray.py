"""Submodule mirror of `autoray.autoray`: the reference writes to `_FUNC_ALIASES`
(desolver/backend/torch_backend.py:56).  Test infrastructure only."""
_FUNC_ALIASES = {}
