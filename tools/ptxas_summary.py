"""Per-kernel register / spill / stack summary of the ptxas logs written by lcgp_b200/csrc/Makefile."""
import glob, os, re, subprocess, sys
d = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(__file__), '..', 'lcgp_b200', 'csrc', 'build')
for f in sorted(glob.glob(os.path.join(d, '*.ptxas.log'))):
    txt = open(f).read()
    for m in re.finditer(r"Compiling entry function '([^']+)'.*?\n.*?\n\s*(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers", txt):
        name = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r'\(.*', '', name).replace('void lcgp::', '')
        print(f'{os.path.basename(f)[:-10]:12s} {name[:70]:70s} regs {m.group(5):>3s} stack {m.group(2):>4s} spill {m.group(3)}/{m.group(4)}')
