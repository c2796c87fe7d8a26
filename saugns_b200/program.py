"""sauProgram data model in Python (ctypes), per include/sau_program_abi.h.

Two uses:
  * `dump(ptr)` turns any sauProgram (e.g. one built by the reference front
    end) into plain Python data, for tests and debugging;
  * `ProgramBuilder` emits the same event / operator data the reference's
    parseconv would (sau/parser/parseconv.h:282-331,472-517) for programmatic
    workloads -- the synthetic many-voice benchmark programs -- so the
    bench's product arm needs no script front end.  tests/test_program_builder.py
    checks builder output field by field against the reference parser's
    output for the same script.
"""
import ctypes as C
import struct

# ---- enums (sau/program.h, sau/line.h, sau/wave.h) -------------------------
WAVES = ["sin", "tri", "srs", "sqr", "ean", "cat", "eto", "par", "mto", "saw", "hsi", "spa"]
LINES = ["cos", "lin", "sah", "exp", "log", "xpe", "lge", "sqe", "cub", "smo", "ncl", "nhl", "uwh"]
NOISES = ["wh", "gw", "bw", "tw", "re", "vi", "bv"]
POPT_AMP, POPT_NOISE, POPT_WAVE, POPT_RASEG = range(4)
POPP_TIME, POPP_MODE, POPP_PHASE, POPP_SEED = 1, 2, 4, 8
LINEP_STATE, LINEP_STATE_RATIO, LINEP_GOAL, LINEP_GOAL_RATIO = 1, 2, 4, 8
LINEP_TYPE, LINEP_TIME, LINEP_TIME_IF_NEW = 16, 32, 64
TIMEP_SET, TIMEP_DEFAULT, TIMEP_IMPLICIT = 1, 2, 4
POP_USES = ["carr", "camod", "amod", "ramod", "fmod", "rfmod", "pmod", "apmod", "fpmod"]
PVO_NO_ID = 0xFFFF
PMODE_AMP_DIV_VOICES = 1


class Line(C.Structure):
    _fields_ = [("v0", C.c_float), ("vt", C.c_float), ("pos", C.c_uint32), ("end", C.c_uint32),
                ("time_ms", C.c_uint32), ("type", C.c_uint8), ("flags", C.c_uint8)]


class Time(C.Structure):
    _fields_ = [("v_ms", C.c_uint32), ("flags", C.c_uint8)]


class RasOpt(C.Structure):
    """sauRasOpt (sau/program.h:126-132): uint8 line, then flags:10, func:6, level:8 packed by gcc
    into the same 32-bit unit (bits 8.., 18.., 24..); stated as the raw word, ctypes would start a
    new unit for the bit-fields."""
    _fields_ = [("w0", C.c_uint32), ("alpha", C.c_uint32)]

    @property
    def line(self):
        return self.w0 & 0xff

    @property
    def flags(self):
        return (self.w0 >> 8) & 0x3ff

    @property
    def func(self):
        return (self.w0 >> 18) & 0x3f

    @property
    def level(self):
        return (self.w0 >> 24) & 0xff

    @staticmethod
    def pack(line, flags, func, level):
        return (line & 0xff) | ((flags & 0x3ff) << 8) | ((func & 0x3f) << 18) | ((level & 0xff) << 24)


class Mode(C.Union):
    _fields_ = [("main", C.c_uint8), ("ras", RasOpt)]


class OpData(C.Structure):
    _fields_ = [("id", C.c_uint32), ("params", C.c_uint32), ("time", Time),
                ("pan", C.POINTER(Line)), ("amp", C.POINTER(Line)), ("amp2", C.POINTER(Line)),
                ("freq", C.POINTER(Line)), ("freq2", C.POINTER(Line)), ("pm_a", C.POINTER(Line)),
                ("phase", C.c_uint32), ("seed", C.c_uint32), ("use_type", C.c_uint8),
                ("type", C.c_uint8), ("mode", Mode),
                ("camods", C.c_void_p), ("amods", C.c_void_p), ("ramods", C.c_void_p),
                ("fmods", C.c_void_p), ("rfmods", C.c_void_p), ("pmods", C.c_void_p),
                ("apmods", C.c_void_p), ("fpmods", C.c_void_p)]


class Event(C.Structure):
    _fields_ = [("wait_ms", C.c_uint32), ("vo_id", C.c_uint16), ("carr_op_id", C.c_uint32),
                ("op_count", C.c_uint32), ("op_data_count", C.c_uint32),
                ("op_list", C.c_void_p), ("op_data", C.POINTER(OpData))]


class Program(C.Structure):
    _fields_ = [("events", C.POINTER(Event)), ("ev_count", C.c_size_t), ("mode", C.c_uint16),
                ("vo_count", C.c_uint16), ("op_count", C.c_uint32), ("op_nest_depth", C.c_uint8),
                ("duration_ms", C.c_uint32), ("ampmult", C.c_float), ("name", C.c_char_p),
                ("mp", C.c_void_p), ("parse", C.c_void_p)]


MOD_FIELDS = ["camods", "amods", "ramods", "fmods", "rfmods", "pmods", "apmods", "fpmods"]
LINE_FIELDS = ["pan", "amp", "amp2", "freq", "freq2", "pm_a"]


def _fbits(x):
    return struct.unpack("<I", struct.pack("<f", x))[0]


def _idarr(ptr):
    if not ptr:
        return None
    n = C.c_uint32.from_address(ptr).value
    return list((C.c_uint32 * n).from_address(ptr + 4))


def _line(lp):
    if not lp:
        return None
    l = lp.contents
    return {"v0": _fbits(l.v0), "vt": _fbits(l.vt), "time_ms": l.time_ms, "type": l.type,
            "flags": l.flags}


def dump(ptr):
    """Plain-data view of a sauProgram at address `ptr` (floats as raw bits)."""
    p = Program.from_address(ptr)
    out = {"mode": p.mode, "vo_count": p.vo_count, "op_count": p.op_count,
           "op_nest_depth": p.op_nest_depth, "duration_ms": p.duration_ms,
           "ampmult": _fbits(p.ampmult), "events": []}
    for i in range(p.ev_count):
        e = p.events[i]
        ev = {"wait_ms": e.wait_ms, "vo_id": e.vo_id, "carr_op_id": e.carr_op_id, "ops": []}
        for k in range(e.op_data_count):
            od = e.op_data[k]
            d = {"id": od.id, "params": od.params, "time": (od.time.v_ms, od.time.flags),
                 "phase": od.phase, "seed": od.seed, "use_type": od.use_type, "type": od.type}
            if od.type == POPT_RASEG:
                r = od.mode.ras
                d["mode"] = ("ras", r.line, r.flags, r.func, r.level, r.alpha)
            else:
                d["mode"] = ("main", od.mode.main)
            for f in LINE_FIELDS:
                d[f] = _line(getattr(od, f))
            for f in MOD_FIELDS:
                d[f] = _idarr(getattr(od, f))
            ev["ops"].append(d)
        out["events"].append(ev)
    return out


# ---- builder ----------------------------------------------------------------

def value(v0, goal=None, line="lin", time_ms=None, ratio=False, goal_ratio=None):
    """A parameter value with an optional sweep (README.SAU "Value sweep")."""
    return {"v0": v0, "goal": goal, "line": line, "time_ms": time_ms, "ratio": ratio,
            "goal_ratio": ratio if goal_ratio is None else goal_ratio}


class ProgramBuilder:
    """Builds a sauProgram in ctypes memory: one event per top-level voice at
    t=0 (wait_ms 0), every operator new in its event -- the shape of the
    synthetic benchmark scripts."""

    def __init__(self, ampmult=1.0, amp_div_voices=False, name=b"builder", default_time_ms=1000):
        self.default_time_ms = default_time_ms     # "S t" default, README.SAU
        self.ampmult = ampmult
        self.amp_div_voices = amp_div_voices
        self.name = name
        self._keep = []
        self._events = []       # (wait_ms, [opdata dicts in emit order], carr_id, dur_ms)
        self._next_op = 0
        self._depth = 0

    # -- operator descriptions (nested dicts) --
    @staticmethod
    def wave(wave="sin", freq=None, amp=None, time_ms=None, pan=None, phase=0, mods=None,
             amp2=None, freq2=None, pm_a=None, raw_mods=None):
        """raw_mods: {use: [operator ids]} appended to the lists as given -- ids of operators
        emitted elsewhere (ids count up in visiting order, parent before its modulators); the
        way to state graphs the script language cannot, e.g. a circular reference."""
        return {"type": POPT_WAVE, "mode": WAVES.index(wave), "freq": freq, "amp": amp,
                "time_ms": time_ms, "pan": pan, "phase": phase, "mods": mods or {},
                "amp2": amp2, "freq2": freq2, "pm_a": pm_a, "raw_mods": raw_mods or {}}

    # -- the other operator types (sau/parser.c:1157-1172, 1601-1680, 1990-2000) --
    RAS_FUNCS = "ugbtfa"
    RAS_OPTS = {"p": 1, "h": 2, "z": 4, "s": 8, "v": 16}
    RAS_O_LINE_SET, RAS_O_FUNC_SET = 1 << 6, 1 << 7

    @staticmethod
    def noise(noise="wh", amp=None, time_ms=None, pan=None, mods=None, amp2=None):
        """`N<noise>`: the parser gives every operator the default 440 Hz frequency line."""
        return {"type": POPT_NOISE, "mode": NOISES.index(noise), "freq": 440.0, "amp": amp,
                "time_ms": time_ms, "pan": pan, "phase": 0, "mods": mods or {}, "amp2": amp2,
                "freq2": None, "pm_a": None, "raw_mods": {}, "seeded": True}

    @staticmethod
    def raseg(line="lin", mode="", freq=None, amp=None, time_ms=None, pan=None, mods=None,
              amp2=None, freq2=None, pm_a=None):
        """`R<line> m<mode>`: mode = one function letter of "ugbtfa" and / or option letters
        "hpsvz" (parse_op_mode, sau/parser.c:1601-1680)."""
        flags, func = ProgramBuilder.RAS_O_LINE_SET, 0
        for ch in mode:
            if ch in ProgramBuilder.RAS_FUNCS:
                func = ProgramBuilder.RAS_FUNCS.index(ch)
                flags |= ProgramBuilder.RAS_O_FUNC_SET
            else:
                flags |= ProgramBuilder.RAS_OPTS[ch]
        return {"type": POPT_RASEG, "mode": RasOpt.pack(LINES.index(line), flags, func, 0), "freq": freq,
                "amp": amp, "time_ms": time_ms, "pan": pan, "phase": 0, "mods": mods or {},
                "amp2": amp2, "freq2": freq2, "pm_a": pm_a, "raw_mods": {}, "seeded": True}

    def _next_seed(self):
        """Seeds of N and R operators in deterministic mode: sau_rand32 = SplitMix32 from state 0,
        one draw per seeded operator in script order (sau/parser.c:1169-1170, sau/math.h:329-353)."""
        m = 0xffffffff
        self._seed_pos = (getattr(self, "_seed_pos", 0) + 0x9e3779b9) & m
        z = self._seed_pos
        z = ((z ^ (z >> 16)) * 0x21f0aaad) & m
        z = ((z ^ (z >> 15)) * 0xf35a2d97) & m
        return z ^ (z >> 15)

    def _line(self, spec, default_time_ms, sub=False):
        """sauLine as the parser leaves it for a parameter of a NEW operator:
        STATE|TIME|TIME_IF_NEW with the operator's default time, TYPE unless it
        is a second (".r") value; a sweep adds GOAL and its own shape/time."""
        if spec is None:
            return None
        if not isinstance(spec, dict):
            spec = value(spec)
        l = Line()
        l.v0 = spec["v0"]
        flags = LINEP_STATE | LINEP_TIME | LINEP_TIME_IF_NEW
        if not sub:
            flags |= LINEP_TYPE
        if spec["ratio"]:
            flags |= LINEP_STATE_RATIO
        l.type = LINES.index("lin")
        l.time_ms = default_time_ms
        if spec["goal"] is not None:
            l.vt = spec["goal"]
            flags |= LINEP_GOAL | LINEP_TYPE
            if spec["goal_ratio"]:
                flags |= LINEP_GOAL_RATIO
            l.type = LINES.index(spec["line"])
            if spec["time_ms"] is not None:
                l.time_ms = spec["time_ms"]
                flags &= ~LINEP_TIME_IF_NEW
        l.flags = flags
        self._keep.append(l)
        return C.pointer(l)

    def _idarr(self, ids):
        buf = (C.c_uint32 * (1 + len(ids)))(len(ids), *ids)
        self._keep.append(buf)
        return C.addressof(buf)

    def _emit(self, node, use, level, out, dur_ms):
        """Children first (parseconv.h:352-360), then this operator's data."""
        self._depth = max(self._depth, level)
        op_id = self._next_op
        self._next_op += 1
        seed = self._next_seed() if node.get("seeded") else 0
        # ids are allocated in visiting order: parent before its modulators
        # (sauOpAlloc_update runs before the recursion, parseconv.h:350-356)
        mod_ids = {}
        for use_name, lst in node["mods"].items():     # script order
            if not lst:
                continue
            ids = []
            for child in lst:
                ids.append(self._emit(child, use_name, level + 1, out, dur_ms))
            mod_ids[use_name] = ids
        for use_name, ids in node.get("raw_mods", {}).items():
            mod_ids[use_name] = mod_ids.get(use_name, []) + list(ids)
        implicit = node["time_ms"] is None and level > 0
        t_ms = node["time_ms"] if node["time_ms"] is not None else (
            self.default_time_ms if level > 0 else dur_ms)
        od = {"id": op_id, "node": node, "use": POP_USES.index(use), "mods": mod_ids,
              "implicit": implicit, "t_ms": t_ms, "seed": seed}
        out.append(od)
        return op_id

    def add_voice(self, carrier, wait_ms=0, vo_id=None):
        """wait_ms: time since the event before (a `|` in a script: the duration of what came before);
        vo_id: the voice slot (the converter gives a voice that has ended its slot back: the voices of
        `A | B | C` all live in slot 0, sau/parser/parseconv.h voice allocation); default: a slot of its own."""
        ops = []
        dur = carrier["time_ms"]
        carr_id = self._emit(carrier, "carr", 0, ops, dur)
        self._events.append((wait_ms, ops, carr_id, dur, len(self._events) if vo_id is None else vo_id))

    def finish(self):
        """-> object with .ptr (address of the sauProgram) keeping memory alive."""
        nev = len(self._events)
        evs = (Event * max(nev, 1))()
        total_ms = 0
        start_ms = 0
        vo_count = 0
        for vi, (wait_ms, ops, carr_id, dur, vo_id) in enumerate(self._events):
            start_ms += wait_ms
            vo_count = max(vo_count, vo_id + 1)
            ods = (OpData * len(ops))()
            for k, od in enumerate(ops):
                node = od["node"]
                o = ods[k]
                o.id = od["id"]
                o.params = POPP_TIME | POPP_MODE | POPP_PHASE | POPP_SEED   # new op: parser.c:984-990
                if od["implicit"]:
                    o.time = Time(od["t_ms"], TIMEP_SET | TIMEP_DEFAULT | TIMEP_IMPLICIT)
                else:
                    o.time = Time(od["t_ms"], TIMEP_SET)
                o.use_type = od["use"]
                o.type = node["type"]
                if node["type"] == POPT_RASEG:
                    o.mode.ras.w0 = node["mode"]
                else:
                    o.mode.main = node["mode"]
                o.phase = node["phase"]
                o.seed = od["seed"]
                dflt = od["t_ms"]
                o.amp = self._line(node["amp"] if node["amp"] is not None else 1.0, dflt) or None
                o.freq = self._line(node["freq"], dflt) or None
                if node["pan"] is not None or od["use"] == 0:
                    o.pan = self._line(node["pan"] if node["pan"] is not None else 0.0, dflt) or None
                for name in ("amp2", "freq2"):
                    if node[name] is not None:
                        setattr(o, name, self._line(node[name], dflt, sub=True))
                if node["pm_a"] is not None:
                    o.pm_a = self._line(node["pm_a"], dflt, sub=True)
                for use_name, ids in od["mods"].items():
                    setattr(o, use_name + "s", self._idarr(ids))
            self._keep.append(ods)
            e = evs[vi]
            e.wait_ms = wait_ms
            e.vo_id = vo_id
            e.carr_op_id = carr_id
            e.op_count = 0
            e.op_data_count = len(ops)
            e.op_list = None
            e.op_data = C.cast(ods, C.POINTER(OpData))
            total_ms = max(total_ms, start_ms + dur)
        prg = Program()
        prg.events = C.cast(evs, C.POINTER(Event))
        prg.ev_count = nev
        prg.mode = PMODE_AMP_DIV_VOICES if self.amp_div_voices else 0
        prg.vo_count = vo_count
        prg.op_count = self._next_op
        prg.op_nest_depth = self._depth
        prg.duration_ms = total_ms
        prg.ampmult = self.ampmult
        prg.name = self.name
        self._keep += [evs, prg]
        return BuiltProgram(prg, self._keep)


class BuiltProgram:
    def __init__(self, prg, keep):
        self._prg = prg
        self._keep = keep
        self.ptr = C.addressof(prg)
        self.vo_count = prg.vo_count
        self.op_count = prg.op_count
        self.ev_count = prg.ev_count
        self.duration_ms = prg.duration_ms
