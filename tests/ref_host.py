"""Literal Python restatement of the reference's host-side steps around the hot path -- TEST INFRASTRUCTURE.

  read_fasta      inputs/FASTAPointer.java:67-149 + inputs/Fasta.java:21-39 (gapsRemoved = false)
  build_jplace    core/algos/PlacementProcess.java:568-629 (duplicates), :797-806 (unplaced), :974-1047 (rows)
  java_number     Float.toString / Double.toString layout (what json-simple prints)
"""
import re

import numpy as np


def read_fasta(text: str):
    """[(header, sequence)] : header = first line without '>', sequence = other lines joined, trimmed."""
    records, cur = [], None
    for line in re.split(r"\r\n|\n|\r", text):      # BufferedReader.readLine
        if line == "" or line.startswith("#"):      # :82-87
            continue
        if line[0] == ">":                           # :91-110
            if cur is not None:
                records.append(cur)
            cur = [line[1:], []]
            continue
        cur[1].append(line)
    if cur is not None:
        records.append(cur)
    out = []
    for h, lines in records:
        seq = "".join(lines)
        # String.trim(): code points <= U+0020 off both ends (:143-145)
        b, e = 0, len(seq)
        while b < e and ord(seq[b]) <= 0x20:
            b += 1
        while e > b and ord(seq[e - 1]) <= 0x20:
            e -= 1
        out.append((h, seq[b:e]))
    return out


def java_number(v, as_float=False) -> str:
    x = np.float32(v) if as_float else np.float64(v)
    if np.isnan(x) or np.isinf(x):
        return "null"
    if x == 0:
        return "-0.0" if np.signbit(x) else "0.0"
    s = np.format_float_scientific(x, unique=True, trim="-", exp_digits=1)   # shortest round-trip digits
    mant, exp = s.split("e")
    neg = mant.startswith("-")
    digits = mant.lstrip("-").replace(".", "")
    e = int(exp)
    if -3 <= e < 7:
        if e >= 0:
            ip = (digits + "0" * (e + 1))[:e + 1]
            fp = digits[e + 1:] or "0"
            body = ip + "." + fp
        else:
            body = "0." + "0" * (-e - 1) + digits
    else:
        body = digits[0] + "." + (digits[1:] or "0") + "E" + str(e)
    return ("-" if neg else "") + body


def build_jplace(records, place_one, edge_id, branch_len, guppy=False):
    """records: [(header, seq)]; place_one(seq) -> (status, rows[(node, score f32, lwr f64)]).
    Returns (placements list as python objects with numbers as java strings, not_placed headers)."""
    registered = {}
    placements, not_placed = [], []
    for header, seq in records:                      # :568
        key = seq.replace("-", "")                   # getSequence(true), :593
        sub = header.split(" ")[0] if " " in header else header   # :596-601
        if key in registered:                        # :603-624
            registered[key]["nm"].append([sub, 1])
            continue
        status, rows = place_one(seq)
        if status == 1:                              # :797-806
            not_placed.append(header)
            continue
        assert status == 0
        if not rows:                                 # below nsBound (:974)
            continue
        p = []
        for node, score, lwr in rows:
            distal = np.float32(branch_len[node]) / np.float32(2)
            if guppy:                                # :1005-1016
                p.append([java_number(distal, True), int(edge_id[node]), java_number(lwr), java_number(score, True), "0.0"])
            else:                                    # :1017-1024
                p.append([int(edge_id[node]), java_number(score, True), java_number(lwr), java_number(distal, True), "0.0"])
        pl = {"p": p, "nm": [[header, 1]]}           # :1036-1047
        placements.append(pl)
        registered[key] = pl
    return placements, not_placed
