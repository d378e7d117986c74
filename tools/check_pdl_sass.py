"""Build check: no global-memory load may be scheduled before griddepcontrol.wait (SASS ACQBULK) in a kernel that is
launched with programmatic stream serialization (the compiler hoists loads through `const __restrict__` pointers above
the asm barrier).  Usage: python tools/check_pdl_sass.py [lib.so]; exit code 1 lists the offending kernels."""
import os, re, subprocess, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(REPO, "crowdnav_prediction_attngraph_b200", "libcrowdnav_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
bad, n = [], 0
for block in out.split("Function : ")[1:]:
    name = block.split("\n", 1)[0].strip()
    ins = re.findall(r"/\*[0-9a-f]{4,}\*/\s+(.*?);", block)
    if not any(i.strip().startswith("ACQBULK") for i in ins):
        continue
    n += 1
    for i in ins:
        op = i.strip().split()[0] if not i.strip().startswith("@") else i.strip().split()[1]
        if op.startswith("ACQBULK"):
            break
        if re.match(r"(LDG|LD\.|LD$|LDGSTS|ATOMG|ATOM\.|ATOM$|RED\.|RED$|STG|ST\.|ST$)", op):
            bad.append((name, i.strip()))
print("%d kernels with griddepcontrol.wait checked" % n)
for name, i in bad:
    print("EARLY GLOBAL ACCESS  %s : %s" % (name, i))
sys.exit(1 if bad else 0)
