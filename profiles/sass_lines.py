"""Static SASS instruction count per source line of one kernel (code size, not execution counts):
    cuobjdump -xelf all libcz_b200.so && nvdisasm -g cz_kernels.sm_100a.cubin > disasm.txt
    python profiles/sass_lines.py disasm.txt '<mangled kernel name substring>' [top_n]
Instruction fetch bounds the dynamics kernels, so this is the map of where the code bytes go."""
import re
import sys
from collections import defaultdict


def main(path, kernel, top=40):
    cur, inside = None, False
    per_line, per_file = defaultdict(int), defaultdict(int)
    total = 0
    for l in open(path):
        if l.startswith("//-----") and ".text." in l:
            inside = kernel in l
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,5}\*/", l):
            per_line[cur] += 1
            per_file[cur[0] if cur else "?"] += 1
            total += 1
    print("total SASS instructions", total, dict(per_file))
    for (f, ln), c in sorted(per_line.items(), key=lambda kv: -kv[1])[:top]:
        print(f"{c:6d}  {f}:{ln}")
    return per_line


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 40)
