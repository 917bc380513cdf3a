import sys
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import numpy as NP, torch
from prisim_b200 import synthetic as S
from prisim_b200.interferometry import InterferometerArray, SimpleTime
cfg = S.config2()
method = sys.argv[1] if len(sys.argv) > 1 else "auto"
for lst in (0.0, 40.0, 170.0):
    ia = InterferometerArray(cfg["labels"], cfg["baselines"], cfg["channels"], telescope=cfg["telescope"], latitude=cfg["latitude"], skycoords="radec", pointing_coords="hadec", device=0)
    ia.audit_baselines = 128
    ia.skyvis_method = method
    ia.audit_tolerance = 1.0
    ia.observe(SimpleTime(2451545.0, lst), {"Tnet": 300.0}, NP.ones(1024), cfg["pointing_hadec"], cfg["skymodel"], cfg["t_acc"])
    print(method, lst, ia.precision_report)
