#!/usr/bin/env python3
"""API wrapper: python api_pinn_newton_raphson.py input.json output.json (see pinn_fem_b200/_api_pinn.py)."""
import sys
from pathlib import Path

_ROOT = Path(__file__).resolve().parents[1]
if str(_ROOT) not in sys.path:
    sys.path.insert(0, str(_ROOT))

from pinn_fem_b200._api_pinn import main  # noqa: E402

if __name__ == "__main__":
    main("nr", "api_pinn_newton_raphson.py")
