"""296 airfoil layups (one wave of two CTAs per SM) through gxb_section_properties: the command under ncu for the section-kernel capture."""
import sys
sys.path.insert(0, "/root/repo")
from gxbeam_b200 import api, workloads
w = workloads.airfoil_layup_batch(296)
sec = api.Section(w["nn"], w["conn"])
r = sec.properties(w["nodes"], w["materials"], w["theta"])
print("ok", bool(r["ok"].all()), "device ms", r["device_ms"])
sec.close()
