"""TEST-ONLY: build tests/hostcheck/libathena_b200_emu.so -- the product's device sources
(athena-gamma_b200/csrc) compiled for the host against tests/hostcheck/emu/cuda_runtime.h.

The only source transformation is the launch syntax, which g++ cannot parse:
    kernel<T...><<<grid, block, shmem, stream>>>(args)
 -> ab_emu::launch(grid, block, COOP, [&]() { kernel<T...>(args); })
Everything else (kernels, launch geometry, work lists, the mesh / task glue, the C ABI) is the
product's code unchanged.  -ffp-contract=off mirrors nvcc -fmad=false.
"""
import os
import re
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "athena-gamma_b200", "csrc")
GEN = os.path.join(HERE, "_gen", "pkg", "csrc")
SO = os.path.join(HERE, "libathena_b200_emu.so")
SOURCES = ["ab_kernels.cu", "ab_flux_nu.cu", "ab_mesh.cu", "ab_smr.cpp", "ab_smr_kernels.cu"]
# kernels that use __syncthreads / warp shuffles (run with one fiber per thread)
COOP = {"k_cons2prim", "k_new_dt", "k_history", "k_flux_ppm_x1", "k_flux_ppm_t", "k_mesh_new_dt"}


def _balanced_back(s, end, open_c, close_c):
    """s[end-1] == close_c: index of the matching open_c"""
    depth = 0
    i = end - 1
    while i >= 0:
        if s[i] == close_c:
            depth += 1
        elif s[i] == open_c:
            depth -= 1
            if depth == 0:
                return i
        i -= 1
    raise ValueError("unbalanced")


def _balanced_fwd(s, start, open_c, close_c):
    depth = 0
    i = start
    while i < len(s):
        if s[i] == open_c:
            depth += 1
        elif s[i] == close_c:
            depth -= 1
            if depth == 0:
                return i
        i += 1
    raise ValueError("unbalanced")


def _split_top(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    out.append(cur.strip())
    return out


def rewrite_launches(src):
    out, pos, n = "", 0, 0
    while True:
        k = src.find("<<<", pos)
        if k < 0:
            break
        ns = k
        if src[ns - 1] == ">":
            ns = _balanced_back(src, ns, "<", ">")
        m = re.search(r"[A-Za-z_][A-Za-z_0-9:]*$", src[:ns])
        name_start = m.start()
        name = src[name_start:k]
        ce = src.index(">>>", k)
        cfg = _split_top(src[k + 3:ce])
        assert src[ce + 3] == "(", src[ce:ce + 40]
        ae = _balanced_fwd(src, ce + 3, "(", ")")
        args = src[ce + 4:ae]
        coop = "true" if m.group(0) in COOP else "false"
        out += src[pos:name_start]
        out += "ab_emu::launch(%s, %s, %s, [&]() { %s(%s); })" % (cfg[0], cfg[1], coop, name, args)
        pos = ae + 1
        n += 1
    return out + src[pos:], n


def needs_build(SO=SO):
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)
            if f.endswith((".cu", ".cuh", ".h", ".cpp"))]
    deps += [os.path.join(HERE, "emu", "cuda_runtime.h"), os.path.abspath(__file__),
             os.path.join(ROOT, "include", "athena_b200.h")]
    return any(os.path.getmtime(f) > t for f in deps)


def build(force=False, asan=False):
    """asan: a second library with AddressSanitizer (out-of-bounds accesses of "device" memory
    in any kernel / copy); load it with LD_PRELOAD=$(gcc -print-file-name=libasan.so)"""
    if asan:
        return _build(SO[:-3] + "_asan.so", ["-fsanitize=address", "-fno-omit-frame-pointer"], force)
    return _build(SO, [], force)


def _build(SO, extra, force):
    if not force and not needs_build(SO):
        return SO
    os.makedirs(GEN, exist_ok=True)
    total = 0
    objs = []
    procs = []
    for f in sorted(os.listdir(CSRC)):
        if not f.endswith((".cu", ".cuh", ".h", ".cpp")):
            continue
        text = open(os.path.join(CSRC, f)).read()
        text, n = rewrite_launches(text)
        total += n
        # generated files keep their names so that the relative #includes resolve inside _gen/
        with open(os.path.join(GEN, f), "w") as fh:
            fh.write(text)
    # the sources include "../../include/athena_b200.h" relative to csrc/
    inc = os.path.join(HERE, "_gen", "include")
    os.makedirs(inc, exist_ok=True)
    shutil.copy(os.path.join(ROOT, "include", "athena_b200.h"), inc)
    for f in SOURCES:
        o = os.path.join(GEN, f + (".asan.o" if extra else ".o"))
        cmd = ["g++", "-O1", "-g", "-std=c++17", "-fPIC", "-ffp-contract=off", "-x", "c++",
               "-I", os.path.join(HERE, "emu"), "-I", GEN, "-include", "cuda_runtime.h",
               "-DAB_HOST_EMU=1", "-Wno-unknown-pragmas"] + extra + ["-c", os.path.join(GEN, f), "-o", o]
        procs.append((f, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                                          text=True)))
        objs.append(o)
    for f, p in procs:
        log = p.communicate()[0]
        if p.returncode != 0:
            raise RuntimeError("g++ failed on %s:\n%s" % (f, log[-6000:]))
    subprocess.run(["g++", "-shared", "-o", SO] + extra + objs + ["-ldl"], check=True)
    return SO


if __name__ == "__main__":
    import sys
    print(build(force=True, asan="--asan" in sys.argv))
