"""DFMA rate with 0 / 1 / 2 independent integer-pipe instructions per DFMA (gb200_fp64_issue_probe)."""
import ctypes as C, sys
sys.path.insert(0, ".")
import gradus_b200 as gb
from gradus_b200 import _cabi as cabi
ens = gb.EnsembleB200(devices=(0,))
lib = cabi.load()
for mix in (0, 1, 2, 3, 4, 5):
    out = C.c_double()
    cabi.check(lib.gb200_fp64_issue_probe(ens.ctx(0), mix, C.byref(out)), ens.ctx(0))
    print(f"mode {mix}: {out.value:.2f} T(FP64 instr x 2)/s   [0: DFMA x const; 1, 2: + 1, 2 LOP3 per DFMA; 3: DFMA with three register operands; 4: DMUL / DADD with register operands; 5: three register operands, one shared by consecutive DFMAs]")
