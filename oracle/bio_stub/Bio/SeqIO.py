"""Bio.SeqIO.parse(path, "fasta") as scripts/generate_kmers.py uses it: records with .name (first word of the header),
.seq (the sequence lines joined, case kept) and len().  Test infrastructure only (see Bio/__init__.py)."""


class _Record:
    def __init__(self, name, seq):
        self.name = name
        self.id = name
        self.seq = seq

    def __len__(self):
        return len(self.seq)


def parse(path, fmt):
    assert fmt == "fasta"
    name, parts = None, []
    with open(path) as f:
        for line in f:
            if line.startswith(">"):
                if name is not None:
                    yield _Record(name, "".join(parts))
                words = line[1:].split()
                name, parts = (words[0] if words else ""), []
            elif name is not None:
                parts.append("".join(line.split()))
    if name is not None:
        yield _Record(name, "".join(parts))
