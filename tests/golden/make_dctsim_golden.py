"""Golden outputs of the UNMODIFIED reference src/dct-sim.py (run in the build container, where /root/reference exists):

    python tests/golden/make_dctsim_golden.py

  example-dbsearch.txt     dct-sim.py --dct example-dct.npz --db example-dct.npz --top 3 --threshold 0.3
  example-dbsearch-g6pd.txt  dct-sim.py --dct G6PD-dct.npz --db example-dct.npz            (defaults: --top 5 --threshold 0.25)
  example-allsim.txt       dct-sim.py --dct example-dct.npz                               (all-vs-all)
Only the result files are kept (the "dct loaded ..." / timing lines on stdout are not part of them)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = '/root/reference/src/dct-sim.py'


def run(out, *args):
    path = os.path.join(HERE, out)
    if os.path.exists(path):
        os.remove(path)          # the reference appends
    subprocess.run([sys.executable, REF, *args, '--output', path], check=True, cwd=HERE, stdout=subprocess.DEVNULL)
    print(out, sum(1 for _ in open(path)), 'lines')


if __name__ == '__main__':
    run('example-dbsearch.txt', '--dct', 'example-dct.npz', '--db', 'example-dct.npz', '--top', '3', '--threshold', '0.3')
    run('example-dbsearch-g6pd.txt', '--dct', 'G6PD-dct.npz', '--db', 'example-dct.npz')
    run('example-allsim.txt', '--dct', 'example-dct.npz')
