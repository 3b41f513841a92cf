"""Import shim for `xformers` (reference pins 0.0.26.post1, requirements.txt:10). With
XFORMERS_DISABLED=true the reference selects its own BasicSelfAttention (attention.py:158-161),
so these symbols only need to exist. TEST INFRASTRUCTURE ONLY."""
