"""A/B library builds on the c3 varlen + name-stage workload. usage: probe_ab3.py a.so b.so ..."""
import os, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for so in sys.argv[1:] * 2:
    print(so, flush=True)
    subprocess.run([sys.executable, os.path.join(root, "scripts", "probe3.py")], env=dict(os.environ, CLK_LIB=os.path.join(root, so), GEN="1000000", REPS="20"), check=False)
