"""tests/emul/prep.py — TEST INFRASTRUCTURE.  Rewrites the CUDA launch syntax of a .cu file so that the file compiles as plain C++
against tests/emul/cuda_runtime.h (the lock-step warp model):

    kernel<<<grid, block[, smem[, stream]]>>>(args);   ->   emu::launch("kernel", dim3(grid), dim3(block), smem, [&]{ kernel(args); });
    extern __shared__ T name[];                        ->   T *name = (T *)emu::dyn_smem();

Nothing else is touched: the kernels, the device functions they call and the host layer are compiled as they are."""
import re
import sys


def split_top(s: str):
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


def close_paren(s: str, i: int) -> int:
    """index of the ')' matching the '(' at s[i]"""
    depth = 0
    for j in range(i, len(s)):
        if s[j] == "(":
            depth += 1
        elif s[j] == ")":
            depth -= 1
            if depth == 0:
                return j
    raise ValueError("unbalanced launch")


def rewrite(src: str) -> str:
    out, at = "", 0
    for m in re.finditer(r"(\w+)<<<(.*?)>>>\s*\(", src, re.S):
        if m.start() < at:
            continue
        end = close_paren(src, m.end() - 1)
        cfg = split_top(m.group(2))
        args = src[m.end():end]
        out += src[at:m.start()]
        smem = cfg[2] if len(cfg) > 2 else "0"
        out += f'emu::launch("{m.group(1)}", dim3({cfg[0]}), dim3({cfg[1]}), {smem}, [&]{{ {m.group(1)}({args}); }})'
        at = end + 1
    out += src[at:]
    # dynamic shared memory: `extern __shared__ T name[];` -> a pointer into the block's buffer
    out = re.sub(r"extern\s+__shared__\s+__align__\(\d+\)\s+unsigned char\s+(\w+)\s*\[\s*\]\s*;", r"unsigned char *\1 = (unsigned char *)emu::dyn_smem();", out)
    return re.sub(r"extern\s+__shared__\s+([\w:]+)\s+(\w+)\s*\[\s*\]\s*;", r"\1 *\2 = (\1 *)emu::dyn_smem();", out)


def no_ptx(src: str) -> str:
    """inline PTX has no meaning here: the prefetch hint of ns_model.cuh becomes a no-op (its address computations stay)"""
    src = src.replace('asm volatile("prefetch.global.L1 [%0];" :: "l"(p));', "(void)p;")
    src = src.replace('asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));', "t = 0;")          # (profiling builds only)
    # lane 0 acting for the warp: rendezvous around the block (cuda_runtime.h, L0Guard)
    return src.replace("#define L0 if (c.scalar || c.lane == 0)", "#define L0 EMU_L0")


if __name__ == "__main__":
    import os
    src = open(sys.argv[1]).read()
    root = os.path.abspath(os.path.join(os.path.dirname(os.path.abspath(sys.argv[1])), "..", ".."))
    res = no_ptx(rewrite(src)).replace('"../../include/', f'"{root}/include/')
    assert "<<<" not in res
    open(sys.argv[2], "w").write(f'#line 1 "{sys.argv[1]}"\n' + res)
