"""Runs the REFERENCE'S OWN source for the hot path under Python 3 -- TEST INFRASTRUCTURE ONLY.

The reference (/root/reference/script/*.py) is Python 2.7 and imports h5py / pywrapfst / magphase at module
level, so it cannot be imported here.  Its arithmetic for this path, however, is plain numpy + scipy code.
This module reads the reference files where they lie (nothing is copied into the repo), applies a MECHANICAL
Python 2 -> 3 transform, and executes

  * whole modules whose imports resolve:  const, segmentaxis, speech_manip, matrix_operations, label_manip,
    data_manipulation, util, fst_functions_wrapped (with `pywrapfst` -> oracle/minifst.py);
  * selected METHODS of `class Synthesiser` from synth_simple.py / synth_halfphone.py, lifted by name into a bare
    class (`RefSimple`, `RefHalfphone`) whose attributes the caller sets the way `Synthesiser.__init__` would
    from the voice file (synth_simple.py:76-139, synth_halfphone.py:172-300).

The transform (see `py2to3`) is syntactic only:
  print statements (incl. `print >>f, x` and trailing commas) -> print();  `raise E, "msg"` -> raise E("msg");
  `except E, e` -> `except E as e`;  tuple parameters `def f(self, (a, b))` -> unpacked in the body;
  every `/` -> a helper that floors when both operands are integers (Python 2's `/`);
  builtins `zip / map / filter / range` return lists, `xrange = range` (Python 2's builtins).
No numeric expression is rewritten.  One behaviour is NOT reproduced by executing the source: under Python 2
`'%s' % numpy.float64` prints 12 significant digits (numpy < 1.14) where Python 3 prints the shortest
round-trip form; minifst.Compiler.py2_str re-parses weights through '%.12g' to model that.

Only tests/ (and the fixture generator tests/golden/make_reference_fixtures.py) use this module; it needs
/root/reference and therefore never runs on the GPU box -- the fixtures it produced are committed.
"""
from __future__ import annotations

import ast
import builtins
import io
import os
import re
import sys
import tokenize
import types

REF_ROOT = os.environ.get("SNICKERY_REFERENCE", "/root/reference")
SCRIPT = os.path.join(REF_ROOT, "script")


def available():
    return os.path.isfile(os.path.join(SCRIPT, "synth_halfphone.py"))


# ----------------------------------------------------------------------------- source transform
def _convert_prints(src):
    """Rewrites Python 2 print STATEMENTS as print() calls, token based (strings and comments are never touched)."""
    lines = src.split("\n")
    try:
        toks = list(tokenize.generate_tokens(io.StringIO(src).readline))
    except (tokenize.TokenError, IndentationError):
        toks = None
    if toks is None:
        raise SyntaxError("reference source does not tokenize")
    edits = []   # (start (row, col), end (row, col), replacement)
    n = len(toks)
    i = 0
    stmt_start = True
    depth = 0
    while i < n:
        t = toks[i]
        if t.type == tokenize.OP and t.string in "([{":
            depth += 1
        elif t.type == tokenize.OP and t.string in ")]}":
            depth -= 1
        if stmt_start and t.type == tokenize.NAME and t.string == "print" and depth == 0:
            # find the end of this simple statement: NEWLINE, or ';' at depth 0
            j = i + 1
            d = 0
            while j < n:
                tj = toks[j]
                if tj.type == tokenize.OP and tj.string in "([{":
                    d += 1
                elif tj.type == tokenize.OP and tj.string in ")]}":
                    d -= 1
                elif d == 0 and (tj.type == tokenize.NEWLINE or (tj.type == tokenize.OP and tj.string == ";")):
                    break
                elif d == 0 and tj.type == tokenize.COMMENT:
                    break
                j += 1
            args = toks[i + 1:j]
            args = [a for a in args if a.type not in (tokenize.NL,)]
            start = t.start
            end = toks[j].start if j < n else t.end
            if not args:
                edits.append((start, end, "print()"))
            else:
                a0, a1 = args[0].start, args[-1].end
                text = _slice(lines, a0, a1)
                if args[0].type == tokenize.OP and args[0].string == ">>":
                    # print >> f, x, y
                    rest = _slice(lines, args[1].start, a1)
                    parts = _split_top_commas(rest)
                    f = parts[0]
                    body = ", ".join(p for p in parts[1:] if p.strip())
                    trailing = len(parts) > 1 and parts[-1].strip() == ""
                    call = "print(%s%sfile=%s%s)" % (body, ", " if body else "", f.strip(), ", end=' '" if trailing else "")
                    edits.append((start, end, call))
                else:
                    parts = _split_top_commas(text)
                    trailing = parts[-1].strip() == ""
                    body = ", ".join(p for p in parts if p.strip())
                    # `print (x)` / `print(x)` with a single parenthesised argument is already a valid call
                    call = "print(%s%s)" % (body, ", end=' '" if trailing else "")
                    edits.append((start, end, call))
            i = j
            stmt_start = False
            continue
        if t.type in (tokenize.NEWLINE, tokenize.INDENT, tokenize.DEDENT, tokenize.ENCODING, tokenize.NL, tokenize.COMMENT):
            stmt_start = True if t.type != tokenize.COMMENT else stmt_start
        elif t.type == tokenize.OP and t.string in (":", ";") and depth == 0:
            stmt_start = True
        else:
            stmt_start = False
        i += 1
    # apply edits from the end
    for (start, end, rep) in sorted(edits, reverse=True):
        (r0, c0), (r1, c1) = start, end
        if r0 == r1:
            lines[r0 - 1] = lines[r0 - 1][:c0] + rep + lines[r0 - 1][c1:]
        else:
            lines[r0 - 1] = lines[r0 - 1][:c0] + rep + lines[r1 - 1][c1:]
            del lines[r0:r1]
    return "\n".join(lines)


def _slice(lines, a, b):
    (r0, c0), (r1, c1) = a, b
    if r0 == r1:
        return lines[r0 - 1][c0:c1]
    out = [lines[r0 - 1][c0:]] + lines[r0:r1 - 1] + [lines[r1 - 1][:c1]]
    return " ".join(s.rstrip("\\").strip() for s in out)


def _split_top_commas(text):
    parts, depth, cur, q = [], 0, [], None
    i = 0
    while i < len(text):
        ch = text[i]
        if q:
            cur.append(ch)
            if ch == "\\":
                i += 1
                if i < len(text):
                    cur.append(text[i])
            elif text.startswith(q, i):
                cur.extend(q[1:])
                i += len(q) - 1
                q = None
        elif ch in "'\"":
            q = text[i:i + 3] if text[i:i + 3] in ("'''", '"""') else ch
            cur.extend(q)
            i += len(q) - 1
        elif ch in "([{":
            depth += 1
            cur.append(ch)
        elif ch in ")]}":
            depth -= 1
            cur.append(ch)
        elif ch == "," and depth == 0:
            parts.append("".join(cur))
            cur = []
        else:
            cur.append(ch)
        i += 1
    parts.append("".join(cur))
    return parts


_RAISE = re.compile(r"^(\s*)raise\s+([A-Za-z_][\w.]*)\s*,\s*(.+)$")
_EXCEPT = re.compile(r"^(\s*except\s+[^:,]+?)\s*,\s*([A-Za-z_]\w*)\s*:")
_TUPLE_PARAM = re.compile(r"^(\s*)def\s+(\w+)\s*\((.*)\)\s*:\s*$")


def _fix_lines(src):
    out = []
    lines = src.split("\n")
    i = 0
    while i < len(lines):
        line = lines[i]
        m = _RAISE.match(line)
        if m and not line.lstrip().startswith("#"):
            # the message may continue over backslash-continued lines
            msg = m.group(3)
            while msg.rstrip().endswith("\\") and i + 1 < len(lines):
                i += 1
                msg = msg.rstrip()[:-1] + " " + lines[i].strip()
            line = "%sraise %s(%s)" % (m.group(1), m.group(2), msg)
        m = _EXCEPT.match(line)
        if m:
            line = "%s as %s:%s" % (m.group(1), m.group(2), line[m.end():])
        m = _TUPLE_PARAM.match(line)
        if m and "(" in m.group(3):
            params = _split_top_commas(m.group(3))
            unpack = []
            for k, p in enumerate(params):
                ps = p.strip()
                if ps.startswith("(") and ps.endswith(")"):
                    name = "_tuple_arg_%d" % k
                    unpack.append("%s = %s" % (ps, name))
                    params[k] = " " + name
            if unpack:
                line = "%sdef %s(%s):" % (m.group(1), m.group(2), ",".join(params).strip())
                out.append(line)
                indent = m.group(1) + "    "
                out.extend(indent + u for u in unpack)
                i += 1
                continue
        out.append(line)
        i += 1
    return "\n".join(out)


class _Py2Division(ast.NodeTransformer):
    """a / b  ->  _py2_div(a, b);  a /= b  ->  a = _py2_div(a, b)"""

    def visit_BinOp(self, node):
        self.generic_visit(node)
        if isinstance(node.op, ast.Div):
            return ast.copy_location(ast.Call(func=ast.Name(id="_py2_div", ctx=ast.Load()), args=[node.left, node.right],
                                              keywords=[]), node)
        return node

    def visit_AugAssign(self, node):
        self.generic_visit(node)
        if isinstance(node.op, ast.Div):
            load = ast.parse(ast.unparse(node.target), mode="eval").body
            call = ast.Call(func=ast.Name(id="_py2_div", ctx=ast.Load()), args=[load, node.value], keywords=[])
            return ast.copy_location(ast.Assign(targets=[node.target], value=call), node)
        return node


def _py2_div(a, b):
    import numpy as np
    ints = (int, np.integer)
    if isinstance(a, ints) and isinstance(b, ints) and not isinstance(a, bool) and not isinstance(b, bool):
        return a // b
    if isinstance(a, np.ndarray) and isinstance(b, (np.ndarray,) + ints) and a.dtype.kind in "iu" and \
            (not isinstance(b, np.ndarray) or b.dtype.kind in "iu"):
        return a // b
    return a / b


def py2to3(src, filename="<reference>"):
    """Python 2 source text -> compiled Python 3 code object with Python 2's `/`."""
    src = src.replace("\t", "        ")
    src = _fix_lines(src)
    src = _convert_prints(src)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", SyntaxWarning)      # '\.' in non-raw regex strings of the reference
        tree = ast.parse(src, filename=filename)
    tree = _Py2Division().visit(tree)
    ast.fix_missing_locations(tree)
    return compile(tree, filename, "exec")


def _py2_builtins():
    return {
        "_py2_div": _py2_div,
        "zip": lambda *a: list(builtins.zip(*a)),
        "map": lambda f, *a: list(builtins.map(f, *a)),
        "filter": lambda f, a: list(builtins.filter(f, a)),
        "range": lambda *a: list(builtins.range(*a)),
        "xrange": builtins.range,
        "unicode": str,
        "basestring": str,
        "long": int,
        "raw_input": input,
        "reduce": __import__("functools").reduce,
    }


# ----------------------------------------------------------------------------- module loading
_MODULES = {}
# third-party modules the reference imports that are absent here and irrelevant to the arithmetic
_STUB_NAMES = ("h5py", "pylab", "magphase", "libutils", "soundfile", "matplotlib", "regex", "pywrapfst", "openfst")


def _stub(name):
    if name in ("pywrapfst", "openfst"):
        from oracle import minifst
        return minifst
    m = types.ModuleType(name)
    m.__dict__["__getattr__"] = lambda attr: (_ for _ in ()).throw(
        AttributeError("%s.%s: stubbed third-party module of the reference" % (name, attr)))
    return m


def load_module(name):
    """Executes /root/reference/script/<name>.py (transformed) as a module.  Imports of sibling reference modules
    resolve to modules loaded the same way; absent third-party packages resolve to stubs (pywrapfst -> minifst)."""
    if name in _MODULES:
        return _MODULES[name]
    path = os.path.join(SCRIPT, name + ".py")
    with open(path) as f:
        code = py2to3(f.read(), path)
    mod = types.ModuleType("snickery_ref_" + name)
    mod.__file__ = path
    mod.__dict__.update(_py2_builtins())
    real_import = builtins.__import__

    def ref_import(modname, globals=None, locals=None, fromlist=(), level=0):
        top = modname.split(".")[0]
        if level == 0 and os.path.isfile(os.path.join(SCRIPT, top + ".py")) and top not in sys.builtin_module_names:
            return load_module(top)
        if level == 0 and top in _STUB_NAMES:
            try:
                return real_import(modname, globals, locals, fromlist, level)
            except ImportError:
                return _stub(top)
        return real_import(modname, globals, locals, fromlist, level)

    b = dict(vars(builtins))
    b["__import__"] = ref_import
    mod.__dict__["__builtins__"] = b
    mod.__dict__["__name__"] = "snickery_ref_" + name     # `if __name__ == '__main__'` blocks do not run
    _MODULES[name] = mod
    try:
        exec(code, mod.__dict__)
    except BaseException:
        del _MODULES[name]
        raise
    return mod


# ----------------------------------------------------------------------------- method lifting
def extract_methods(filename, class_name, names):
    """Source text of the named methods of `class class_name` in a reference file (the LAST definition of each, as
    Python would bind it), dedented to one class level.  Comment lines at any indent stay with the method."""
    path = os.path.join(SCRIPT, filename)
    with open(path) as f:
        text = f.read().replace("\t", "        ")
    lines = text.split("\n")
    # lines that lie inside a multi-line string literal are not structure (synth_halfphone.py:980-1009 holds a
    # column-0 `def` inside a docstring): blank them in the copy used for scanning
    scan = list(lines)
    for tok in tokenize.generate_tokens(io.StringIO(text).readline):
        if tok.type == tokenize.STRING and tok.end[0] > tok.start[0]:
            for r in range(tok.start[0], tok.end[0]):
                scan[r] = "#"
    lines, real = scan, lines
    start = None
    for i, l in enumerate(lines):
        if re.match(r"class\s+%s\b" % re.escape(class_name), l):
            start = i
            break
    if start is None:
        raise KeyError("class %s not found in %s" % (class_name, filename))
    end = len(lines)
    for i in range(start + 1, len(lines)):
        l = lines[i]
        if l.strip() and not l.lstrip().startswith("#") and not l[0].isspace():
            end = i
            break
    found = {}
    i = start + 1
    while i < end:
        m = re.match(r"    def\s+(\w+)\s*\(", lines[i])
        if not m:
            i += 1
            continue
        j = i + 1
        # a (possibly multi-line) signature, then the body: everything up to the next statement at class level
        while j < end:
            l = lines[j]
            s = l.strip()
            if s and not s.startswith("#") and (len(l) - len(l.lstrip())) <= 4 and not _inside_continuation(lines, i, j):
                break
            j += 1
        found[m.group(1)] = (i, j)
        i = j
    out = []
    for n in names:
        if n not in found:
            raise KeyError("method %s not found in %s:%s" % (n, filename, class_name))
        a, b = found[n]
        out.append("\n".join(real[a:b]))
    return out


def _inside_continuation(lines, a, j):
    """True if line j continues an open bracket / backslash from the lines a..j-1 (cheap bracket count)."""
    depth = 0
    for l in lines[a:j]:
        code = l.split("#", 1)[0]
        depth += sum(code.count(c) for c in "([{") - sum(code.count(c) for c in ")]}")
    return depth > 0 or lines[j - 1].rstrip().endswith("\\")


def build_class(class_name, filename, ref_class, method_names, namespace):
    """A bare class made of reference methods.  `namespace`: the module-level names those methods use."""
    srcs = extract_methods(filename, ref_class, method_names)
    src = "class %s(object):\n%s\n" % (class_name, "\n\n".join(srcs))
    code = py2to3(src, os.path.join(SCRIPT, filename))
    ns = dict(_py2_builtins())
    ns.update(namespace)
    exec(code, ns)
    return ns[class_name]


# ----------------------------------------------------------------------------- the two reference synthesisers
_SIMPLE_METHODS = ["get_tree_for_greedy_search", "set_join_weights", "set_target_weights", "greedy_joint_search",
                   "get_selection_vector", "truncate_join_streams", "truncate_target_streams", "report", "start_clock",
                   "stop_clock"]
_HALFPHONE_METHODS = ["get_tree_for_greedy_search", "set_join_weights", "set_target_weights", "preselect_units_quinphone",
                      "preselect_units_acoustic", "preselect_units_monophone_then_acoustic", "viterbi_search",
                      "greedy_joint_search", "report", "start_clock", "stop_clock", "get_target_scores_per_stream",
                      "get_join_scores_per_stream", "get_natural_distance_vectorised", "aggregate_squared_errors_by_stream",
                      "make_on_the_fly_join_lattice_BLOCK_DIRECT"]


def _common_namespace():
    import math
    import pickle
    import timeit

    import numpy as np
    import scipy.spatial
    fstw = load_module("fst_functions_wrapped")
    ns = {
        "np": np, "numpy": np, "scipy": scipy, "math": math, "sys": sys, "os": os, "timeit": timeit, "pickle": pickle,
        "const": load_module("const"), "segment_axis": load_module("segmentaxis").segment_axis,
        "weight": load_module("speech_manip").weight, "break_quinphone": load_module("label_manip").break_quinphone,
        "VERY_BIG_WEIGHT_VALUE": load_module("const").VERY_BIG_WEIGHT_VALUE,
        "make_target_sausage_lattice": fstw.make_target_sausage_lattice,
        "cost_cache_to_compiled_fst": fstw.cost_cache_to_compiled_fst,
        "get_best_path_SIMP": fstw.get_best_path_SIMP, "get_shortest_path": fstw.get_shortest_path,
        "get_data_dump_name": lambda config, **kw: "/nonexistent/voice",
    }
    return ns


def _init_common(self, config, F, Jc):
    """What Synthesiser.__init__ sets before the weighting calls (synth_simple.py:60-133; synth_halfphone.py:160-270)."""
    import numpy as np
    self.config = dict(config)
    self.config.setdefault("join_cost_type", "natural2")    # a required key of every halfphone config (config/*.cfg)
    self.verbose = False
    self.mode_of_operation = "normal"
    self.stream_list_target = config["stream_list_target"]
    self.stream_list_join = config["stream_list_join"]
    self.datadims_target = config["datadims_target"]
    self.datadims_join = config["datadims_join"]
    self.target_representation = config.get("target_representation", "epoch")
    self.holdout_samples = 0
    self.train_unit_features_unweighted = np.asarray(F, dtype=np.float32)     # HDF5 dtype 'f' (train_simple.py:137)
    self.join_contexts_unweighted = np.asarray(Jc, dtype=np.float32)
    self.number_of_units = self.train_unit_features_unweighted.shape[0]
    jcw = config["join_cost_weight"]
    # APPLY_JCW_ON_TOP (synth_simple.py:44,128-133; synth_halfphone.py:260-270)
    self.set_target_weights(np.array(config["target_stream_weights"]) * (1.0 - jcw))
    self.set_join_weights(np.array(config["join_stream_weights"]) * jcw)


_CLASSES = {}


def RefSimple(config, F, Jc):
    """synth_simple.Synthesiser restricted to the search path; methods are the reference's own code."""
    if "simple" not in _CLASSES:
        _CLASSES["simple"] = build_class("RefSimple", "synth_simple.py", "Synthesiser", _SIMPLE_METHODS, _common_namespace())
    self = _CLASSES["simple"]()
    _init_common(self, config, F, Jc)
    if "truncate_target_streams" in config:
        self.truncate_target_streams(config["truncate_target_streams"])
    if "truncate_join_streams" in config:
        self.truncate_join_streams(config["truncate_join_streams"])
    return self


def RefHalfphone(config, F, Jc, train_unit_names=None):
    """synth_halfphone.Synthesiser restricted to the search path.  The KD-trees are built with the constructor calls of
    synth_halfphone.py:379 / :385-402 (those lines live inside __init__, between voice loading and GUI set-up, and
    are restated here; everything else below is lifted code)."""
    import numpy as np
    import scipy.spatial
    if "halfphone" not in _CLASSES:
        ns = _common_namespace()
        ns["make_synthesis_condition_name"] = None
        _CLASSES["halfphone"] = build_class("RefHalfphone", "synth_halfphone.py", "Synthesiser", _HALFPHONE_METHODS, ns)
    cls = _CLASSES["halfphone"]
    self = cls()
    self.make_synthesis_condition_name = lambda: "cond"
    _init_common(self, config, F, Jc)
    self.train_unit_names = train_unit_names
    method = config.get("preselection_method", "quinphone")
    if config.get("greedy_search", False) and self.target_representation == "epoch":
        self.get_tree_for_greedy_search()
    elif method == "acoustic":
        self.tree = scipy.spatial.cKDTree(self.train_unit_features, leafsize=100, compact_nodes=False, balanced_tree=False)
    elif method == "monophone_then_acoustic":
        const = load_module("const")
        self.phonetrees, self.phonetrees_index_converters = {}, {}
        monophones = np.array([q.split(const.label_delimiter)[2] for q in train_unit_names])
        for phone in dict(zip(monophones, monophones)):
            train = self.train_unit_features[monophones == phone, :]
            self.phonetrees[phone] = scipy.spatial.cKDTree(train, leafsize=10, compact_nodes=False, balanced_tree=False)
            self.phonetrees_index_converters[phone] = np.arange(self.number_of_units)[monophones == phone]
    if train_unit_names is not None:
        # unit index for quinphone preselection (synth_halfphone.py:281-292)
        bq = load_module("label_manip").break_quinphone
        self.unit_index = {}
        for (i, quinphone) in enumerate(train_unit_names):
            for form in bq(quinphone):
                self.unit_index.setdefault(form, []).append(i)
    return self


def last_viterbi_cost():
    """Float32 path weight of the most recent minifst.shortestpath call (instrumentation of the stand-in)."""
    from oracle import minifst
    return getattr(minifst, "_last_path_weight", None)
