"""TEST INFRASTRUCTURE ONLY -- empty stand-in so reference src/train.py:8 imports (plots are out of scope)."""
