"""tests/golden/multimat_tinyobj.json: what the reference's vendored tinyobjloader reports for multimat.obj
(per-face material ids, the MTL materials).  Needs oracle/_ref/ref_objmat (make -C oracle ref)."""
import json
import os
import shutil
import subprocess
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
d = tempfile.mkdtemp()
os.makedirs(os.path.join(d, "models", "materials"))
shutil.copy(os.path.join(HERE, "multimat.obj"), os.path.join(d, "models"))
shutil.copy(os.path.join(HERE, "multimat.mtl"), os.path.join(d, "models", "materials"))
out = subprocess.run([os.path.join(ROOT, "oracle", "_ref", "ref_objmat"), os.path.join(d, "models", "multimat.obj"),
                      os.path.join(d, "models", "materials")], capture_output=True, text=True, check=True).stdout
json.dump(json.loads(out), open(os.path.join(HERE, "multimat_tinyobj.json"), "w"), indent=1)
print(out)
