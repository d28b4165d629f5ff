import sys, os, json, subprocess, tempfile, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import pathlib
from test_host_cpp import edit_deck, BOUNDED, LHC, load_dump
exe = os.path.join(ROOT, "uintah_b200", "host", "rmcrt_test")
rad = ("</RMCRT>", "<Radiometer> <viewAngle> 60 </viewAngle> <normalVector> [0,0,2] </normalVector> <locationsMin> [0.4,0.4,0.0] </locationsMin>"
                   " <locationsMax> [0.5,0.5,0.06] </locationsMax> <nRays> 30 </nRays> </Radiometer>\n</RMCRT>")
for name, edits in (("none", ()), ("bounded", (BOUNDED,)), ("lhc", (LHC,)), ("rad", (rad,)), ("bounded+rad", (BOUNDED, rad)), ("all", (BOUNDED, LHC, rad))):
    for env in ({}, {"RMCRT_PAGEABLE_HOST": "1"}):
        tmp = pathlib.Path(tempfile.mkdtemp())
        deck = edit_deck(tmp, "bm1_1L_double.ups", *edits)
        r = subprocess.run([exe, deck, "--dump", str(tmp / "d"), "--knob", "skipZeroWeightLegs=0"], capture_output=True, text=True, env={**os.environ, **env})
        out = load_dump(str(tmp / "d"))
        print(name, "pageable" if env else "pinned", "rc", r.returncode, "max|divQ|", float(np.abs(out["divQ"]).max()), "volq", float(np.abs(out["radiationVolq"]).max()), flush=True)
