"""SASS evidence per kernel family from the in-tree library (no GPU needed):
    python tools/sass_evidence.py > profiles/rNN_sass_excerpts.txt
Counts the Blackwell mnemonics per tcgen05 kernel and prints short neighbourhoods of the interesting instructions."""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "clg_vqa_b200", "lib", "libclgvqa.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)
MN = ["UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "STTM", "UTCBAR", "SYNCS", "ELECT", "MUFU.EX2"]
rows = []
for f in funcs[1:]:
    name = f.split("\n", 1)[0].strip()
    dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    n_inst = len(re.findall(r"^\s+/\*[0-9a-f]{4,6}\*/\s+\S", f, flags=re.M))
    rows.append((dem, n_inst, {m: len(re.findall(r"\b" + re.escape(m) + r"\b", f)) for m in MN}, f))
fam = ("gemm_tc_kernel", "attn_pipe_fwd", "attn_pipe_bwd", "attn_tc_")
print("# SASS evidence per kernel family: cuobjdump -sass of the in-tree clg_vqa_b200/lib/libclgvqa.so (sm_100a)")
print("# UTCHMMA = tcgen05.mma (kind::f16, bf16 operands; .2CTA = cta_group::2), UTMALDG / UTMASTG = TMA tensor load / store,")
print("# LDTM / STTM = tcgen05.ld / tcgen05.st, UTCBAR = tcgen05.commit, SYNCS = mbarrier operations, ELECT = elect.sync,")
print("# MUFU.EX2 = the exponentials of the softmax / GeLU.  Counts are static instruction counts of the kernel body.\n")
print("%-104s %6s %s" % ("kernel (demangled, argument list cut)", "instr", " ".join("%8s" % m for m in MN)))
for dem, n, cnt, _ in sorted((r for r in rows if any(k in r[0] for k in fam)), key=lambda r: r[0]):
    print("%-104s %6d %s" % (re.sub(r"\(.*", "", dem)[:104], n, " ".join("%8d" % cnt[m] for m in MN)))


def excerpt(pattern, mnem, before=5, after=5):
    for dem, n, cnt, body in rows:
        if pattern in dem and cnt.get(mnem, 0):
            lines = [l for l in body.split("\n") if re.search(r"/\*[0-9a-f]{4,6}\*/\s+\S", l)]
            i = next(k for k, l in enumerate(lines) if mnem in l)
            return dem, [re.sub(r"\s+/\* 0x[0-9a-f]+ \*/", "", l).rstrip() for l in lines[max(0, i - before):i + after]]
    return None, []


for pat, mn, title in [
        ("gemm_tc_kernel<256, true, true, 2, 2>", "UTCHMMA", "GEMM main loop: cta_group::2 MMA issued by the elected lane, descriptors in uniform registers"),
        ("gemm_tc_kernel<256, true, true, 2, 2>", "UTMALDG", "GEMM: TMA operand loads (cta_group::2 pair, half of the B tile per CTA)"),
        ("gemm_tc_kernel<256, true, true, 2, 3>", "UTMASTG", "GEMM GeLU-pair epilogue: TMA stores of the output boxes"),
        ("gemm_tc_kernel<256, false, false, 2, 0>", "UTCHMMA", "weight-gradient GEMM (both operands MN-major through the descriptors)"),
        ("attn_pipe_fwd_kernel<5, 3>", "UTCHMMA", "attention forward (S <= 96): S = Q K^T"),
        ("attn_pipe_fwd_wide_kernel<false>", "STTM", "attention forward (128 < S <= 160): exponentials written back to TMEM"),
        ("attn_pipe_fwd_wide_kernel<false>", "UTCHMMA", "attention forward (128 < S <= 160): MMA issue"),
        ("attn_pipe_bwd_kernel<5, false>", "UTCHMMA", "attention backward: transposed products through MN-major descriptors")]:
    dem, ls = excerpt(pat, mn)
    print("\n## %s\n## %s" % (title, re.sub(r"\(.*", "", dem or pat)))
    for l in ls:
        print(l)
