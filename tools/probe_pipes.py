import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import emosaic_b200 as emo
from tools.probe import probe_int_pipe
names = {0: "IMAD", 1: "VABSDIFF4.ACC", 2: "VIMNMX3(+IADD)", 3: "match mix v1", 4: "f16x2 fma (HFMA2/HADD2 by ptxas)", 5: "f16x2 add (HFMA2/HADD2 by ptxas)",
         6: "VABSDIFF4 + f16x2 fma 1:1", 7: "VABSDIFF4 + f16x2 add 1:1", 8: "VABSDIFF4 + IMAD 1:1", 9: "HMNMX2 (+IADD)", 10: "VABSDIFF4 + HMNMX2 1:1"}
for w in range(11):
    v = probe_int_pipe(0, w)
    print(f"{w} {names[w]:40s} {v/1e12:8.2f} T thread-instr/s  = {v/148/1.965e9:6.1f} lanes/clk/SM")
