"""Build the drop-in proof: the reference's OWN callers -- src/kitti.cpp (the KITTI driver) and test/test.cpp (its Boost.Test
suite) -- compiled UNCHANGED from the read-only reference tree against libviso_b200/host (the reference's API over the
C-ABI) and linked with libviso_b200.so.  OpenCV / Boost / Eigen are not installed in this image, so the header
stand-ins of compat/ are on the include path; with the real libraries installed, drop -Icompat.

    python tools/build_dropin.py            -> build/dropin/ref_kitti, build/dropin/ref_tester, build/dropin/fopen_redirect.so

The binaries are git-ignored build products; they travel to the GPU box with the snapshot (the reference tree does
not exist there), where tests/test_host_cpp.py runs them.
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE = os.environ.get("VISO_REFERENCE_DIR", "/root/reference")
OUT = os.path.join(ROOT, "build", "dropin")
CXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"


def targets():
    return [os.path.join(OUT, n) for n in ("ref_kitti", "ref_tester", "fopen_redirect.so")]


def available():
    return os.path.exists(os.path.join(REFERENCE, "src", "kitti.cpp"))


def build(force=False):
    """returns the list of built files, or None when the reference tree is not here (prebuilt files are used then)"""
    if not available():
        return None
    os.makedirs(OUT, exist_ok=True)
    lib_dir = os.path.join(ROOT, "libviso_b200")
    host = os.path.join(lib_dir, "host", "viso.cpp")
    srcs = [host, os.path.join(REFERENCE, "src", "kitti.cpp"), os.path.join(REFERENCE, "test", "test.cpp"), os.path.abspath(__file__)]
    srcs += [os.path.join(lib_dir, "host", h) for h in ("viso.h", "mvg.h", "misc.h", "estimation.h")]
    newest = max(os.path.getmtime(s) for s in srcs)
    if not force and all(os.path.exists(t) and os.path.getmtime(t) >= newest for t in targets()):
        return targets()
    inc = ["-I" + os.path.join(ROOT, "compat")]
    obj = os.path.join(OUT, "host_viso.o")
    subprocess.check_call([CXX, "-std=c++17", "-O1", "-w"] + inc + ["-c", host, "-o", obj])
    link = [obj, "-L" + lib_dir, "-lviso_b200", "-Wl,-rpath,$ORIGIN/../../libviso_b200", "-lz"]
    # the reference's own flags: -std=c++0x, no optimisation level (src/CMakeLists.txt:2)
    subprocess.check_call([CXX, "-std=c++11", "-w"] + inc + [os.path.join(REFERENCE, "src", "kitti.cpp")] + link + ["-o", targets()[0]])
    subprocess.check_call([CXX, "-std=c++11", "-w"] + inc + [os.path.join(REFERENCE, "test", "test.cpp")] + link + ["-o", targets()[1]])
    subprocess.check_call(["gcc", "-shared", "-fPIC", "-O1", os.path.join(ROOT, "tests", "host", "fopen_redirect.c"), "-ldl", "-o", targets()[2]])
    return targets()


if __name__ == "__main__":
    sys.path.insert(0, ROOT)
    from libviso_b200 import build as b
    b.build()
    print(build(force="--force" in sys.argv))
