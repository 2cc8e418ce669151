"""Builds the reference's own CUDA kernels into oracle/_ref/ (see oracle/ref_cuda/Makefile).
Run by `__graft_entry__.build()` when /root/reference is present (the build container); the GPU box
only uses the prebuilt, git-ignored .so files.  Test / bench infrastructure only."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def main(ref="/root/reference/brainevent"):
    if not os.path.isdir(ref):
        sys.stderr.write(f"[build_ref] {ref} not present; keeping whatever oracle/_ref already holds\n")
        return
    subprocess.run(["make", "-C", os.path.join(HERE, "ref_cuda"), f"-j{min(8, os.cpu_count() or 1)}", f"REF={ref}"],
                   check=True, stdout=subprocess.DEVNULL)


if __name__ == "__main__":
    main()
