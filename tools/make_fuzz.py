"""Random (seeded) Cartesian stencils in the inlined, staged form a CUDA-class backend receives:
differential-test fodder for the planner (oracle vs generated kernels, bit for bit).  Writes
stencils/fuzz_<nn>.iir.   python tools/make_fuzz.py [count]

Validity rules the generator keeps (what Dawn's optimizer guarantees for real programs):
  * a stage never reads, at a horizontal offset, a field written in the same stage; a field written by
    a Parallel multistage is not read at a vertical offset inside it
  * accumulated horizontal extents stay within the 3-point halo
  * sequential multistages may read their own output one level behind the sweep
  * every temporary is written before it is read
  * vertical offsets never leave the allocation (k + 1 <= ksize; k - 1 only from the second level on)"""
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from iir_builder import FULL, IIRBuilder, Interval, call, lit, where  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "stencils")


def make(seed):
    rng = random.Random(seed)
    b = IIRBuilder("fuzz_%02d" % seed)
    n_in = rng.randint(2, 4)
    ins = [b.field("in%d" % n) for n in range(n_in)]
    extra = []
    if rng.random() < 0.6:
        extra.append(b.field("vj", "j"))
    if rng.random() < 0.4:
        extra.append(b.field("mij", "ij"))
    if rng.random() < 0.3:
        extra.append(b.field("colk", "k"))
    outs = [b.field("out%d" % n) for n in range(rng.randint(1, 2))]
    g = b.global_var("alpha", "Double", 0.25) if rng.random() < 0.5 else None

    def leaf(readable, reach, allow_k):
        """One access: readable = list of (handle, max horizontal reach, vertical offsets allowed)."""
        h, hreach, kset = rng.choice(readable)
        r = min(reach, hreach)
        di = rng.randint(-r, r) if r else 0
        dj = rng.randint(-r, r) if r else 0
        dk = rng.choice(kset) if (allow_k and kset) else 0
        return h(di, dj, dk)

    def expr(readable, reach, depth, allow_k=True):
        if depth == 0 or rng.random() < 0.25:
            x = leaf(readable, reach, allow_k)
            if rng.random() < 0.2:
                return x * lit(rng.choice([0.5, 2.0, -1.5, 0.125]))
            return x
        a = expr(readable, reach, depth - 1, allow_k)
        c = expr(readable, reach, depth - 1, allow_k)
        kind = rng.random()
        if kind < 0.35:
            return a + c
        if kind < 0.6:
            return a - c
        if kind < 0.8:
            return a * c
        if kind < 0.88:
            return call("gridtools::dawn::math::max", a, c)
        if kind < 0.94 and g is not None:
            return a * g + c
        return where(a > c, a, c)

    n_ms = rng.randint(1, 2)
    ms_list = []
    produced = []        # (handle, reach budget left for readers, vertical offsets) of earlier multistages
    for m in range(n_ms):
        order = rng.choice(["Parallel", "Parallel", "Forward", "Backward"])
        behind = {"Forward": -1, "Backward": 1}.get(order)
        # Parallel stages only look one level up: storages carry ksize + 1 levels like the reference's
        # drivers, so k + 1 is always allocated, while k - 1 at the first level is not
        in_k = [0, 0, 1] if order == "Parallel" else [0, 0, behind]
        this_out = outs[m % len(outs)]
        # stored fields of earlier multistages: centre only (stage extents propagate across
        # multistages) and never the field this multistage writes (a stage computed redundantly in the
        # halo must not read what a later stage of the same kernel writes; Dawn versions such fields)
        readable0 = [(h, 2, in_k) for h in ins] + [(h, 1, []) for h in extra] + \
                    [(h, 0, [0]) for h in produced if h is not this_out]
        stages = []
        tmps = []
        n_st = rng.randint(1, 3)
        for s in range(n_st):
            last = s == n_st - 1
            # readers of a temporary may reach 1 further, inputs 2: keeps the total within 3
            reach = 1
            readable = list(readable0) + [(t, 1, []) for t in tmps]
            stmts = []
            if not last and rng.random() < 0.8:
                t = b.tmp("t%d_%d" % (m, s))
                stmts.append(b.assign(t, expr(readable0 if not tmps else readable, reach, 2,
                                              allow_k=order == "Parallel")))
                new_tmp = t
            else:
                new_tmp = None
            # stored fields are written by the last stage only: an earlier stage has a horizontal extent
            # (its temporaries are read at offsets) and would be computed redundantly in the halo
            if last:
                o = outs[m % len(outs)]
                e = expr(readable, reach, 3)
                if behind is not None and rng.random() < 0.7:
                    e = e + o(0, 0, behind) * lit(0.5)
                if rng.random() < 0.3:
                    if behind is None:
                        stages.append(b.stage(b.do(Interval("Start", 3, 0, 0), *stmts, b.assign(o, e)),
                                              b.do(Interval(3, "End", 1, 0), *stmts,
                                                   b.assign(o, e * lit(2.0)))))
                    else:
                        # the first level of the sweep cannot look behind
                        first = Interval("Start", "Start", 0, 0) if behind == -1 else Interval("End", "End", 0, 0)
                        rest = Interval("Start", "End", 1, 0) if behind == -1 else Interval("Start", "End", 0, -1)
                        e0 = expr(readable, reach, 2, allow_k=False)
                        stages.append(b.stage(b.do(first, *stmts, b.assign(o, e0)),
                                              b.do(rest, *stmts, b.assign(o, e))))
                    if new_tmp is not None:
                        tmps.append(new_tmp)
                    continue
                if behind is not None:
                    first = Interval("Start", "Start", 0, 0) if behind == -1 else Interval("End", "End", 0, 0)
                    rest = Interval("Start", "End", 1, 0) if behind == -1 else Interval("Start", "End", 0, -1)
                    e0 = expr(readable, reach, 2, allow_k=False)
                    stages.append(b.stage(b.do(first, *stmts, b.assign(o, e0)),
                                          b.do(rest, *stmts, b.assign(o, e))))
                else:
                    stages.append(b.stage(b.do(FULL, *stmts, b.assign(o, e))))
            elif stmts:
                stages.append(b.stage(b.do(FULL, *stmts)))
            if new_tmp is not None:
                tmps.append(new_tmp)
        if stages:
            ms_list.append(b.multistage(order, *stages))
        produced.append(outs[m % len(outs)])
    b.stencil(*ms_list)
    return b


def main():
    count = int(sys.argv[1]) if len(sys.argv) > 1 else 12
    os.makedirs(OUT, exist_ok=True)
    for seed in range(count):
        with open(os.path.join(OUT, "fuzz_%02d.iir" % seed), "w") as f:
            f.write(make(seed).dumps())
    print("wrote %d fuzz stencils" % count)


if __name__ == "__main__":
    main()
