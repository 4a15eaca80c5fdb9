"""md5 of the SASS body and instruction count of every kernel in a build of libpicles_b200.so:
    python profiles/sass_hashes.py [lib.so] > profiles/<name>.txt
Used to show that a source change left the measured kernels alone: `r02_sass_hashes_measured.txt` is the build every GPU
number of round 2 was taken with, `r02_sass_hashes_final.txt` the final one (after picles_params_t::nan_eest_rejects was
added): the specialised k_advance instantiations and k_project_remesh have the same hash in both."""
import hashlib
import os
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "picles_b200", "libpicles_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
name, body = None, {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1)
        body[name] = []
        continue
    if name and re.match(r"\s+/\*[0-9a-f]{4,5}\*/", line):
        body[name].append(line.strip())
dem = subprocess.run(["c++filt"], input="\n".join(body), capture_output=True, text=True).stdout.splitlines()
for k, d in sorted(zip(body, dem), key=lambda x: x[1]):
    print(hashlib.md5("\n".join(body[k]).encode()).hexdigest(), "%6d" % len(body[k]), re.sub(r"\((?!anonymous).*", "", d))
