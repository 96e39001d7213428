#!/usr/bin/env python
"""Offline evaluation of a 3DGS experiment directory (the evaluation half of the reference's pretrain_eval_attention.py):
   python tools/eval_pose.py --exp_path <exp> --images <dir> --out results.json      (see 6dgs_b200/eval_driver.py)"""
import importlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
importlib.import_module("6dgs_b200.eval_driver").main()
