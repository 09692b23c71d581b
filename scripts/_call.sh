python scripts/blend_timing.py 16 256 32
python scripts/blend_timing.py 16 512 32
