"""Measured per-class issue peaks (scgpu_int_peak_gops) -> profiles/int_peaks_*.txt.  Run on the GPU box."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import libsafecrypto_b200 as sc
names = {0: "IMAD", 1: "IMAD.HI", 2: "IADD3+LOP", 3: "LOP3", 4: "SHF", 5: "IMAD.WIDE", 6: "IMAD+IADD3", 7: "FFMA", 8: "DFMA",
         9: "mont-bfly(5 ops)", 10: "barrett32 butterflies (IMAD,IMAD.HI,IMAD,2 IADD)", 11: "float-quotient butterflies (FFMA,2 IMAD,2 IADD)"}
for k in range(12):
    print("%s %.1f G/s" % (names[k], sc.int_peak_gops(k, 4096, 0)))
