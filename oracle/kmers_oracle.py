"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's genome-wide guide generation
(/root/reference/scripts/generate_kmers.py:55-136), pinned against the unmodified script's output on tests/golden/kmers.*
(tests/test_kmers.py).  Used as the checker of gsx_generate_kmers on fresh inputs; never imported by the product.

Restated per position instead of per str.find() call:
  pam set      generate_kmers.py:55-68   N's replaced first-N-first, breadth first, in the order A, C, T, G
  + strand     generate_kmers.py:70-100  PAM occurrence at i: k-mer chr[i-k:i] at i-k (PAM at the end) or chr[i+|pam|:...] at i (--start)
  - strand     generate_kmers.py:70-100  occurrences of the reverse-complemented PAMs, k-mer on the other side, reverse complemented
  filters      generate_kmers.py:96-117  position >= 0, full length, ACGT only
  rows         generate_kmers.py:119-136 id = prefix + chr:pos1:sense; records shorter than min_chr_length skipped
"""
COMP = {"A": "T", "T": "A", "C": "G", "G": "C"}


def revcomp(s):
    return "".join(COMP[c] for c in reversed(s))


def pam_set(pam):
    queue = [pam]
    while any("N" in p for p in queue):
        p = queue.pop(0)
        if "N" not in p:
            queue.append(p)
        else:
            at = p.index("N")
            queue.extend(p[:at] + c + p[at + 1:] for c in "ACTG")
    return queue


def records(path):
    name, parts = None, []
    for line in open(path):
        if line.startswith(">"):
            if name is not None:
                yield name, "".join(parts)
            w = line[1:].split()
            name, parts = (w[0] if w else ""), []
        elif name is not None:
            parts.append("".join(line.split()))
    if name is not None:
        yield name, "".join(parts)


def generate(path, pam="NGG", k=20, min_chr_length=0, prefix="", start=False):
    out = ["id,sequence,pam,chromosome,position,sense\n"]
    fw = pam_set(pam)
    for name, seq in records(path):
        if len(seq) < min_chr_length:
            continue
        s = seq.upper()
        for sense, pams in (("+", fw), ("-", [revcomp(p) for p in fw])):
            before = (sense == "+") != bool(start)              # the k-mer lies in front of the PAM occurrence
            for p in pams:
                for i in range(len(s) - len(p) + 1):
                    if not s.startswith(p, i):
                        continue
                    a = i - k if before else i + len(p)
                    if a < 0 or a + k > len(s):
                        continue
                    kmer = s[a:a + k]
                    if any(c not in "ACGT" for c in kmer):
                        continue
                    pos1 = (a if before else i) + 1
                    out.append("%s%s:%d:%s,%s,%s,%s,%d,%s\n" % (prefix, name, pos1, sense, kmer if sense == "+" else revcomp(kmer),
                                                                pam, name, pos1, sense))
    return "".join(out)
