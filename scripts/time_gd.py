import json, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from bench_gd import gd_iterations_per_second
print(json.dumps(gd_iterations_per_second(torch.device("cuda", 0)), indent=1))
