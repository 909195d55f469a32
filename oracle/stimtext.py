"""Oracle: parser / flattener / canonical printer for the Stim-text dialect QUITS emits.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Restates, in plain Python, the
grammar produced by the reference's string emitters -- nothing else is ever
generated on this path:

* ``R`` / ``RX`` (+ ``X_ERROR`` / ``Z_ERROR``)         reference ``src/quits/circuit.py:78-104``
* idle ``DEPOLARIZE1``                                 ``circuit.py:106-124``
* ``H`` (+ ``DEPOLARIZE1``)                            ``circuit.py:126-148``
* ``CX`` (+ ``DEPOLARIZE2``)                           ``circuit.py:158-180``
* ``M`` / ``MX`` (error before)                        ``circuit.py:191-218``
* ``MR`` (``X_ERROR`` before and after)                ``circuit.py:228-252``
* ``DETECTOR rec[-k] ...``                             ``circuit.py:262-269``
* ``OBSERVABLE_INCLUDE(i) rec[-k] ...``                ``circuit.py:271-279``
* ``REPEAT n {`` ... ``}`` and ``TICK``                ``circuit.py:58-76``

``PAULI_CHANNEL_1/2`` (vector error rates) are parsed but rejected by the
samplers/analyser: the reference's circuit-level decode calls
``detector_error_model(decompose_errors=False)`` (``decoder/base.py:151``) without
``approximate_disjoint_errors``, so it is only defined for scalar rates.
"""
from __future__ import annotations

import re
from dataclasses import dataclass, field
from typing import List, Tuple

GATES_1Q = ("R", "RX", "H", "M", "MX", "MR")
NOISE_1Q = ("X_ERROR", "Z_ERROR", "DEPOLARIZE1")
NOISE_2Q = ("DEPOLARIZE2",)
ANNOT = ("DETECTOR", "OBSERVABLE_INCLUDE", "TICK")
KNOWN = set(GATES_1Q) | set(NOISE_1Q) | set(NOISE_2Q) | set(ANNOT) | {"CX", "PAULI_CHANNEL_1", "PAULI_CHANNEL_2"}
MEASURING = ("M", "MX", "MR")

_LINE = re.compile(r"^([A-Z_][A-Z_0-9]*)(?:\(([^)]*)\))?\s*(.*)$")
_REC = re.compile(r"^rec\[-(\d+)\]$")


@dataclass
class Instr:
    """One (unflattened) instruction.  ``targets`` are qubit ids, or rec lookbacks k>0."""
    name: str
    args: Tuple[float, ...] = ()
    targets: List[int] = field(default_factory=list)
    body: "List[Instr] | None" = None     # REPEAT only
    count: int = 0                        # REPEAT only


def parse(text: str) -> List[Instr]:
    """Text -> nested instruction list (REPEAT bodies kept as blocks)."""
    stack: List[List[Instr]] = [[]]
    counts: List[int] = []
    for lineno, raw in enumerate(text.splitlines(), 1):
        line = raw.split("#", 1)[0].strip()
        if not line:
            continue
        if line == "}":
            if len(stack) == 1:
                raise ValueError(f"line {lineno}: unmatched '}}'")
            body = stack.pop()
            stack[-1].append(Instr("REPEAT", body=body, count=counts.pop()))
            continue
        if line.startswith("REPEAT"):
            m = re.match(r"^REPEAT\s+(\d+)\s*\{$", line)
            if not m:
                raise ValueError(f"line {lineno}: bad REPEAT header {raw!r}")
            counts.append(int(m.group(1)))
            stack.append([])
            continue
        m = _LINE.match(line)
        if not m or m.group(1) not in KNOWN:
            raise ValueError(f"line {lineno}: unsupported instruction {raw!r}")
        name, argstr, rest = m.group(1), m.group(2), m.group(3)
        args = tuple(float(a) for a in argstr.split(",")) if argstr not in (None, "") else ()
        toks = rest.split()
        if name in ("DETECTOR", "OBSERVABLE_INCLUDE"):
            targets = []
            for t in toks:
                mm = _REC.match(t)
                if not mm or int(mm.group(1)) == 0:
                    raise ValueError(f"line {lineno}: bad record target {t!r}")
                targets.append(int(mm.group(1)))
        else:
            try:
                targets = [int(t) for t in toks]
            except ValueError:
                raise ValueError(f"line {lineno}: bad qubit target in {raw!r}") from None
            if any(t < 0 for t in targets):
                raise ValueError(f"line {lineno}: negative qubit target")
        if name in ("CX",) + NOISE_2Q + ("PAULI_CHANNEL_2",) and len(targets) % 2:
            raise ValueError(f"line {lineno}: {name} needs an even number of targets")
        stack[-1].append(Instr(name, args, targets))
    if len(stack) != 1:
        raise ValueError("unterminated REPEAT block")
    return stack[0]


@dataclass
class FlatOp:
    """Flattened op.  For DETECTOR / OBSERVABLE_INCLUDE ``targets`` are ABSOLUTE measurement indices."""
    name: str
    arg: float
    targets: List[int]


@dataclass
class FlatCircuit:
    ops: List[FlatOp]
    n_qubits: int
    n_meas: int
    n_det: int
    n_obs: int


def flatten(instrs: List[Instr]) -> FlatCircuit:
    """Unroll REPEAT blocks and resolve ``rec[-k]`` against the running measurement count."""
    ops: List[FlatOp] = []
    state = {"meas": 0, "det": 0, "obs": 0, "nq": 0}

    def walk(block: List[Instr]) -> None:
        for ins in block:
            if ins.name == "REPEAT":
                for _ in range(ins.count):
                    walk(ins.body)
                continue
            if ins.name in ("PAULI_CHANNEL_1", "PAULI_CHANNEL_2"):
                raise NotImplementedError(
                    f"{ins.name}: circuit-level decoding is only defined for scalar error rates "
                    "(reference decoder/base.py:151 does not pass approximate_disjoint_errors)")
            if ins.name == "TICK":
                continue
            if ins.name in ("DETECTOR", "OBSERVABLE_INCLUDE"):
                abs_t = []
                for k in ins.targets:
                    if k > state["meas"]:
                        raise ValueError("rec[-%d] looks back past the start of the record" % k)
                    abs_t.append(state["meas"] - k)
                if ins.name == "DETECTOR":
                    ops.append(FlatOp("DETECTOR", float(state["det"]), abs_t))
                    state["det"] += 1
                else:
                    idx = int(ins.args[0])
                    ops.append(FlatOp("OBSERVABLE_INCLUDE", float(idx), abs_t))
                    state["obs"] = max(state["obs"], idx + 1)
                continue
            arg = ins.args[0] if ins.args else 0.0
            if ins.targets:
                state["nq"] = max(state["nq"], max(ins.targets) + 1)
            ops.append(FlatOp(ins.name, float(arg), list(ins.targets)))
            if ins.name in MEASURING:
                state["meas"] += len(ins.targets)

    walk(instrs)
    return FlatCircuit(ops, state["nq"], state["meas"], state["det"], state["obs"])


def parse_flat(text: str) -> FlatCircuit:
    return flatten(parse(text))


# ----------------------------------------------------------------------------------------------
# Canonical (Stim-style) printing, used only to compare against notebook printouts.
# Stim fuses adjacent instructions that have the same name and arguments (never TICK / DETECTOR /
# OBSERVABLE_INCLUDE / REPEAT) and prints numbers in shortest round-trip form.
# ----------------------------------------------------------------------------------------------
def _num(x: float) -> str:
    return str(int(x)) if float(x).is_integer() else repr(float(x))


def canonical_lines(instrs: List[Instr], indent: str = "") -> List[str]:
    out: List[str] = []
    prev: "Instr | None" = None
    for ins in instrs:
        if ins.name == "REPEAT":
            out.append(f"{indent}REPEAT {ins.count} {{")
            out.extend(canonical_lines(ins.body, indent + "    "))
            out.append(indent + "}")
            prev = None
            continue
        fusable = ins.name not in ANNOT
        if fusable and prev is not None and prev.name == ins.name and prev.args == ins.args:
            tail = " ".join(str(t) for t in ins.targets)
            if tail:
                out[-1] = out[-1] + " " + tail
            continue
        head = ins.name
        if ins.args:
            head += "(" + ", ".join(_num(a) for a in ins.args) + ")"
        if ins.name in ("DETECTOR", "OBSERVABLE_INCLUDE"):
            tail = " ".join(f"rec[-{k}]" for k in ins.targets)
        else:
            tail = " ".join(str(t) for t in ins.targets)
        out.append(indent + (head + " " + tail if tail else head))
        prev = ins if fusable else None
    return out


def canonical_text(text: str) -> str:
    return "\n".join(canonical_lines(parse(text)))


def count_top_level(text: str) -> int:
    """``len(stim.Circuit)``: top-level instructions after fusion; a REPEAT block counts 1."""
    return sum(1 for ln in canonical_lines(parse(text)) if not ln.startswith("    ") and ln != "}")
