import json, os, subprocess, sys
code = r'''
import os, sys, json, torch
sys.path.insert(0, os.getcwd())
from sd_lora_trainer_b200 import ops
from scripts.bench_gemm import graph_time
'''
