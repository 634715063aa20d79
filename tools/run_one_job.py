import sys
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import jobs, harness
from envutil_b200.engine import Engine
e=Engine(0)
j=jobs.JOBS[sys.argv[1]]
out=e.render(j)
print(harness.compare(out, harness.oracle_render(j)))
