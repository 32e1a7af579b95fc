"""One text summary per .ncu-rep for profiles/: for every captured launch the raw-page numbers DESIGN.md cites
(tools/ncu_summary.py) followed by the opcode mix, stall reasons and hottest SASS lines (tools/ncu_hot.py).
    python tools/profile_pack.py gpurun_out/r02_ncu_x.ncu-rep > profiles/r02_ncu_x.txt"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def main(rep):
    print("# %s  (ncu --set full --clock-control none --import-source on; per-launch times are cold-cache and serialised)" % os.path.basename(rep))
    print(subprocess.run([sys.executable, os.path.join(HERE, "ncu_summary.py"), rep], capture_output=True, text=True).stdout)
    n = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout.count('"Kernel Name"')
    for i in range(n):
        print("\n## launch %d: opcode mix / stall reasons / hottest SASS lines" % i)
        print(subprocess.run([sys.executable, os.path.join(HERE, "ncu_hot.py"), rep, str(i), "24"], capture_output=True, text=True).stdout)


if __name__ == "__main__":
    main(sys.argv[1])
