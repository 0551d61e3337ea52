"""ctypes binding of the CPU oracle (oracle/gs_oracle.c).  TEST INFRASTRUCTURE ONLY -- see gs_oracle.h."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libgsoracle.so")
REF_BIN = os.path.join(HERE, "_ref", "guidescan")


class Opts(C.Structure):
    _fields_ = [("mismatches", C.c_int32), ("rna_bulges", C.c_int32), ("dna_bulges", C.c_int32),
                ("threshold", C.c_int32), ("start", C.c_int32), ("format_sam", C.c_int32),
                ("complete", C.c_int32), ("n_alt_pams", C.c_int32), ("max_off_targets", C.c_int64),
                ("alt_pams", C.c_char_p * 16)]


class Hit(C.Structure):
    _fields_ = [("abs_pos", C.c_int64), ("sa_row", C.c_uint64), ("distance", C.c_uint32), ("rna", C.c_uint32),
                ("dna", C.c_uint32), ("index_id", C.c_uint32), ("seq", C.c_char * 48)]


class Counters(C.Structure):
    _fields_ = [("nodes", C.c_uint64), ("pam_nodes", C.c_uint64), ("rank_calls", C.c_uint64),
                ("lf_steps", C.c_uint64), ("hits", C.c_uint64)]


def build(force: bool = False) -> str:
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(os.path.join(HERE, "gs_oracle.c")):
        subprocess.check_call(["make", "-s", "-C", HERE, "-f", os.path.join(HERE, "Makefile")])
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.gso_index_from_fasta.restype = C.c_void_p
        L.gso_index_from_fasta.argtypes = [C.c_char_p]
        L.gso_index_from_text.restype = C.c_void_p
        L.gso_index_from_text.argtypes = [C.c_char_p, C.c_uint64, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_uint64)]
        L.gso_index_from_bwt.restype = C.c_void_p
        L.gso_index_from_bwt.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_uint64)]
        L.gso_index_free.argtypes = [C.c_void_p]
        L.gso_index_n.restype = C.c_uint64
        L.gso_index_n.argtypes = [C.c_void_p]
        for f in (L.gso_rank_bwt,):
            f.restype = C.c_uint64
            f.argtypes = [C.c_void_p, C.c_int, C.c_uint64, C.c_int]
        for f in (L.gso_sa, L.gso_sa_direct):
            f.restype = C.c_uint64
            f.argtypes = [C.c_void_p, C.c_int, C.c_uint64]
        L.gso_bwt.restype = C.c_uint8
        L.gso_bwt.argtypes = [C.c_void_p, C.c_int, C.c_uint64]
        L.gso_C.restype = C.c_uint64
        L.gso_C.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.gso_process_kmer.restype = C.c_void_p
        L.gso_process_kmer.argtypes = [C.c_void_p, C.POINTER(Opts), C.c_char_p, C.c_char_p, C.c_char_p, C.c_int, C.POINTER(Counters)]
        L.gso_enumerate_hits.restype = C.c_int
        L.gso_enumerate_hits.argtypes = [C.c_void_p, C.POINTER(Opts), C.c_char_p, C.c_char_p, C.POINTER(C.POINTER(Hit)),
                                         C.POINTER(C.c_uint64), C.POINTER(C.c_float), C.POINTER(C.c_int), C.POINTER(Counters)]
        L.gso_enumerate_file.restype = C.c_int64
        L.gso_enumerate_file.argtypes = [C.c_void_p, C.POINTER(Opts), C.c_char_p, C.c_char_p, C.c_int, C.POINTER(Counters)]
        L.gso_calculate_cfd.restype = C.c_float
        L.gso_calculate_cfd.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p]
        L.gso_free.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def make_opts(mismatches=3, rna_bulges=0, dna_bulges=0, threshold=-1, start=False, fmt="csv", mode="complete",
              max_off_targets=-1, alt_pams=()) -> Opts:
    o = Opts()
    o.mismatches, o.rna_bulges, o.dna_bulges, o.threshold = mismatches, rna_bulges, dna_bulges, threshold
    o.start, o.format_sam, o.complete = int(bool(start)), int(fmt == "sam"), int(mode == "complete")
    o.max_off_targets = max_off_targets
    o.n_alt_pams = len(alt_pams)
    for i, p in enumerate(alt_pams):
        o.alt_pams[i] = p.encode()
    return o


class Index:
    def __init__(self, fasta: str | None, handle=None):
        self.h = handle if handle is not None else lib().gso_index_from_fasta(fasta.encode())
        if not self.h:
            raise RuntimeError("oracle: cannot index %s" % fasta)

    @classmethod
    def from_bwt(cls, bwt_fwd, sa64_fwd, bwt_rev, sa64_rev, chroms):
        """bwt_*: uint8 numpy arrays (0 in the sentinel row), sa64_*: uint32 SA samples every 64 rows"""
        names = (C.c_char_p * len(chroms))(*[c[0].encode() for c in chroms])
        lens = (C.c_uint64 * len(chroms))(*[int(c[1]) for c in chroms])
        h = lib().gso_index_from_bwt(bwt_fwd.ctypes.data, sa64_fwd.ctypes.data, bwt_rev.ctypes.data, sa64_rev.ctypes.data,
                                     len(bwt_fwd), len(chroms), names, lens)
        return cls(None, handle=h)

    def close(self):
        if self.h:
            lib().gso_index_free(self.h)
            self.h = None

    def __del__(self):
        self.close()

    @property
    def n(self) -> int:
        return lib().gso_index_n(self.h)

    def rank_bwt(self, strand, i, c):
        return lib().gso_rank_bwt(self.h, strand, i, ord(c) if isinstance(c, str) else c)

    def sa(self, strand, row):
        return lib().gso_sa(self.h, strand, row)

    def process_kmer(self, opts: Opts, gid: str, seq: str, pam: str, positive=True, counters: Counters | None = None) -> str:
        p = lib().gso_process_kmer(self.h, C.byref(opts), gid.encode(), seq.encode(), pam.encode(), int(positive),
                                   C.byref(counters) if counters is not None else None)
        s = C.string_at(p).decode()
        lib().gso_free(p)
        return s

    def hits(self, opts: Opts, seq: str, pam: str, counters: Counters | None = None):
        hp = C.POINTER(Hit)()
        n = C.c_uint64()
        spec = C.c_float()
        dropped = C.c_int()
        lib().gso_enumerate_hits(self.h, C.byref(opts), seq.encode(), pam.encode(), C.byref(hp), C.byref(n),
                                 C.byref(spec), C.byref(dropped), C.byref(counters) if counters is not None else None)
        out = [(hp[i].abs_pos, hp[i].sa_row, hp[i].distance, hp[i].rna, hp[i].dna, hp[i].index_id, hp[i].seq.decode())
               for i in range(n.value)]
        lib().gso_free(hp)
        return out, spec.value, bool(dropped.value)

    def enumerate_file(self, opts: Opts, kmers_csv: str, out_path: str, nthreads: int = 1) -> Counters:
        ctr = Counters()
        r = lib().gso_enumerate_file(self.h, C.byref(opts), kmers_csv.encode(), out_path.encode(), nthreads, C.byref(ctr))
        if r < 0:
            raise RuntimeError("oracle: enumerate_file failed (%d)" % r)
        return ctr


def have_ref() -> bool:
    return os.path.exists(REF_BIN)


def ref_index(fasta: str, prefix: str, cwd: str | None = None) -> None:
    """Runs the unmodified reference `guidescan index` (sdsl writes temp files into the CWD)."""
    subprocess.check_call([REF_BIN, "index", "--index", prefix, fasta], cwd=cwd or os.path.dirname(os.path.abspath(prefix)),
                          stdout=subprocess.DEVNULL)


def ref_enumerate(prefix: str, kmers_csv: str, out: str, mismatches=3, rna_bulges=0, dna_bulges=0, threshold=None,
                  start=False, fmt="csv", mode="complete", max_off_targets=None, alt_pams=(), threads=1) -> None:
    cmd = [REF_BIN, "enumerate", prefix, "-f", kmers_csv, "-o", out, "-m", str(mismatches), "--rna-bulges", str(rna_bulges),
           "--dna-bulges", str(dna_bulges), "--format", fmt, "--mode", mode, "-n", str(threads)]
    if threshold is not None:
        cmd += ["-t", str(threshold)]
    if start:
        cmd += ["--start"]
    if max_off_targets is not None:
        cmd += ["--max-off-targets", str(max_off_targets)]
    if alt_pams:
        cmd += ["-a"] + list(alt_pams)
    subprocess.check_call(cmd, stdout=subprocess.DEVNULL)
