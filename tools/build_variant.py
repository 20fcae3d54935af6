import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sys
from zkvm_pairings_b200 import build
name=sys.argv[1]; defs=sys.argv[2:]
print(build.build(force=True, defines=defs, out="/root/repo/build/libzkpair_%s.so"%name, verbose=False))
