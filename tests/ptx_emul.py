"""A small interpreter for the inline-PTX integer sequences of nufhe_b200/csrc/ff.cuh.

The device arithmetic is written as carry-chain PTX (add.cc / subc / mad.lo.cc / madc.hi ...) inside `asm(...)`
blocks, which the host build never compiles (ff.cuh has plain-C twins for g++).  This module parses the asm
text and operand lists straight out of the source file and executes them with Python integers, so that the
CPU-only test-suite can check the ACTUAL device sequences on edge values that random GPU tests would never hit
(elements in [p, 2^64) have probability 2^-32).

PTX semantics modelled (32-bit unsigned, one carry flag CC.CF):
    add.cc / addc[.cc]          d = a + b (+ CF);           CF = carry out
    sub.cc / subc[.cc]          d = a - b (- CF);           CF = borrow out
    mad.lo.cc / madc.lo[.cc]    d = lo(a * b) + c (+ CF);   CF = carry out
    mad.hi.cc / madc.hi[.cc]    d = hi(a * b) + c (+ CF);   CF = carry out
    mul.lo / mul.hi / mul.wide / mad.wide (no flags); mov.b64 {lo, hi}, x and mov.b64 x, {lo, hi}
An instruction without `.cc` leaves CF unchanged.  Reading, in a subtract, a flag written by an add (or the
reverse) raises: ptxas keeps borrows in the inverted sense on the hardware and mixing the two families was
observed to misbehave, so the source must never do it.
"""
import re

M32 = (1 << 32) - 1
M64 = (1 << 64) - 1


class AsmBlock:
    def __init__(self, text, outputs, inputs):
        self.text = text          # PTX source, instructions separated by ';'
        self.outputs = outputs    # [(constraint, c_expression)]
        self.inputs = inputs


def _split_top(s, sep):
    out, depth, cur = [], 0, ''
    for ch in s:
        if ch in '([':
            depth += 1
        elif ch in ')]':
            depth -= 1
        if ch == sep and depth == 0:
            out.append(cur)
            cur = ''
        else:
            cur += ch
    if cur.strip():
        out.append(cur)
    return out


def _operands(s):
    res = []
    for item in _split_top(s, ','):
        m = re.match(r'\s*"([^"]+)"\s*\((.*)\)\s*$', item, re.S)
        if m:
            res.append((m.group(1), m.group(2).strip()))
    return res


def preprocess(source, defines):
    """Resolve `#if NAME == n` (or `#if NAME`) / `#else` / `#endif` groups whose NAME is in `defines` (other directives are kept)."""
    out, stack = [], []
    for line in source.split('\n'):
        m = re.match(r'\s*#\s*if\s+(\w+)\s*(?:==\s*(\d+)\s*)?$', line)
        if m and m.group(1) in defines:
            stack.append(defines[m.group(1)] == int(m.group(2)) if m.group(2) else bool(defines[m.group(1)]))
            continue
        if stack and re.match(r'\s*#\s*(if|ifdef|ifndef)\b', line):
            stack.append(None)                       # unrelated nested group: transparent
            out.append(line)
            continue
        if stack and re.match(r'\s*#\s*else\b', line) and stack[-1] is not None:
            stack[-1] = not stack[-1]
            continue
        if stack and re.match(r'\s*#\s*endif\b', line):
            if stack.pop() is None:
                out.append(line)
            continue
        if all(v is not False for v in stack):
            out.append(line)
    text = '\n'.join(out)
    # function-like macros that build asm text (NB_ASM_*): expand the definitions that survived the #if groups
    for m in re.finditer(r'#define (NB_ASM_\w+)\(([^)]*)\)[ \t]+(.*)', text):
        name, params, body = m.group(1), [q.strip() for q in m.group(2).split(',')], m.group(3)

        def expand(call, params=params, body=body):
            args = [a.strip() for a in _split_top(call.group(1), ',')]
            assert len(args) == len(params), call.group(0)
            res = body
            for prm, arg in zip(params, args):
                res = re.sub(r'\b%s\b' % prm, lambda _m, arg=arg: arg, res)
            return re.sub(r'"\s*"', '', res)          # adjacent literals concatenate
        text = re.sub(r'(?<!#define )\b%s\(((?:[^()]|\([^()]*\))*)\)' % name, expand, text)
    return text


def extract_asm_blocks(source, function):
    """All asm blocks inside the body of `function` (first definition found), in order."""
    m = re.search(r'\b%s\s*\([^)]*\)\s*\{' % re.escape(function), source)
    if not m:
        raise KeyError(function)
    i = m.end()
    depth, j = 1, i
    while depth:
        c = source[j]
        depth += c == '{'
        depth -= c == '}'
        j += 1
    body = source[i:j]
    blocks = []
    for am in re.finditer(r'\basm\s*(?:volatile\s*)?\(', body):
        k = am.end()
        depth, e = 1, k
        while depth:
            c = body[e]
            if c == '"':                              # skip string literals
                e += 1
                while body[e] != '"':
                    e += 2 if body[e] == '\\' else 1
            depth += body[e] == '('
            depth -= body[e] == ')'
            e += 1
        inner = body[k:e - 1]
        # string literals first, then ':'-separated operand sections
        pos, text = 0, ''
        while True:
            mm = re.match(r'\s*(?://[^\n]*\n\s*)*"((?:[^"\\]|\\.)*)"', inner[pos:])
            if not mm:
                break
            text += mm.group(1)
            pos += mm.end()
        rest = re.sub(r'//[^\n]*', '', inner[pos:])
        sections = _split_top(rest, ':')
        sections = [s for s in sections[1:]] if rest.strip().startswith(':') else sections
        outs = _operands(sections[0]) if len(sections) > 0 else []
        ins = _operands(sections[1]) if len(sections) > 1 else []
        text = text.replace('\\n', ' ').replace('\\t', ' ')
        blocks.append(AsmBlock(text, outs, ins))
    return blocks


class Machine:
    def __init__(self):
        self.cf = 0
        self.cf_kind = None

    def _read_cf(self, kind):
        if self.cf_kind is not None and self.cf_kind != kind:
            raise AssertionError('carry flag written by a %s is read by a %s' % (self.cf_kind, kind))
        return self.cf

    def run(self, block, env):
        """env: dict of C names -> ints plus callables used in operand expressions.  Returns env (updated)."""
        ops = block.outputs + block.inputs
        vals, widths = [], []
        for idx, (cons, expr) in enumerate(ops):
            w = 64 if 'l' in cons else 32
            widths.append(w)
            is_out = idx < len(block.outputs)
            if is_out and not cons.startswith('+'):
                vals.append(None)
            else:
                vals.append(eval(expr, {}, env) & ((1 << w) - 1))

        local = {}

        def get(tok):
            tok = tok.strip()
            if tok.startswith('%'):
                v = vals[int(tok[1:])]
                assert v is not None, 'operand %s read before it is written' % tok
                return v
            if tok in local:
                assert local[tok] is not None, 'register %s read before it is written' % tok
                return local[tok]
            return int(tok, 0) & M64

        def put(tok, value, width):
            tok = tok.strip()
            if tok.startswith('%'):
                assert widths[int(tok[1:])] == width, tok
                vals[int(tok[1:])] = value
            else:
                assert tok in local, tok
                local[tok] = value

        text = block.text.replace('{', ' { ').replace('}', ' } ')
        text = re.sub(r'\{\s*(%\d+|\w+)\s*,\s*(%\d+|\w+)\s*\}', r'<\1|\2>', text)      # vector operand {a, b}
        text = text.replace('{', ';').replace('}', ';')                                # scopes: flat
        for ins in [i.strip() for i in text.split(';') if i.strip()]:
            m = re.match(r'\.reg\s+\.(\w+)\s+(.*)$', ins)
            if m:
                for name in m.group(2).split(','):
                    local[name.strip()] = None
                continue
            m = re.match(r'mov\.b64\s+<(.+)\|(.+)>\s*,\s*(.+)$', ins)
            if m:
                v = get(m.group(3))
                put(m.group(1), v & M32, 32)
                put(m.group(2), v >> 32, 32)
                continue
            m = re.match(r'mov\.b64\s+(%\d+|\w+)\s*,\s*<(.+)\|(.+)>$', ins)
            if m:
                put(m.group(1), get(m.group(2)) | (get(m.group(3)) << 32), 64)
                continue
            m = re.match(r'([a-z0-9.]+)\s+(.*)$', ins)
            opc, args = m.group(1), [a.strip() for a in m.group(2).split(',')]
            parts = opc.split('.')
            base, cc = parts[0], 'cc' in parts
            assert parts[-1] in ('u32', 's32'), ins
            d = args[0]
            if base in ('add', 'addc'):
                r = get(args[1]) + get(args[2]) + (self._read_cf('add') if base == 'addc' else 0)
                res, c, kind = r & M32, r >> 32, 'add'
            elif base in ('sub', 'subc'):
                r = get(args[1]) - get(args[2]) - (self._read_cf('sub') if base == 'subc' else 0)
                res, c, kind = r & M32, int(r < 0), 'sub'
            elif base in ('mad', 'madc') and 'wide' not in parts:
                pr = get(args[1]) * get(args[2])
                half = (pr & M32) if 'lo' in parts else (pr >> 32)
                r = half + get(args[3]) + (self._read_cf('add') if base == 'madc' else 0)
                res, c, kind = r & M32, r >> 32, 'add'
            elif base == 'mul' and 'wide' in parts:
                res, c, kind = get(args[1]) * get(args[2]), None, None
            elif base == 'mad' and 'wide' in parts:
                res, c, kind = (get(args[1]) * get(args[2]) + get(args[3])) & M64, None, None
            elif base == 'mul':
                pr = get(args[1]) * get(args[2])
                res, c, kind = ((pr & M32) if 'lo' in parts else (pr >> 32)), None, None
            else:
                raise NotImplementedError(ins)
            if cc:
                assert c is not None and c in (0, 1), ins
                self.cf, self.cf_kind = c, kind
            put(d, res, 64 if 'wide' in parts else 32)
        for idx, (cons, expr) in enumerate(block.outputs):
            assert re.match(r'^[A-Za-z_][A-Za-z0-9_]*$', expr), 'output %r is not a plain variable' % expr
            env[expr] = vals[idx]
        return env
