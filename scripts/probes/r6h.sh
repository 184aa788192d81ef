#!/bin/bash
python scripts/cfg_profile.py 3 1.0 2>&1 | grep -E "bound_|total"
python -m pytest tests -m gpu -q -x 2>&1 | tail -4
