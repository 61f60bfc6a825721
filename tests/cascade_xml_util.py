"""Writes tiny purpose-built cascades in OpenCV's new XML format (black-box probes)."""


def write_cascade(path, w, h, stages, features):
    """stages: [(stage_thr, [(feature_idx, thr, left, right), ...])]; features: [[(x,y,w,h,weight), ...]]"""
    s = ['<?xml version="1.0"?>', '<opencv_storage>',
         '<cascade type_id="opencv-cascade-classifier"><stageType>BOOST</stageType>',
         '<featureType>HAAR</featureType>', f'<height>{h}</height>', f'<width>{w}</width>',
         f'<stageParams><maxWeakCount>{max(len(t) for _, t in stages)}</maxWeakCount></stageParams>',
         '<featureParams><maxCatCount>0</maxCatCount></featureParams>', f'<stageNum>{len(stages)}</stageNum>',
         '<stages>']
    for thr, trees in stages:
        s.append(f'<_><maxWeakCount>{len(trees)}</maxWeakCount><stageThreshold>{float(thr)!r}</stageThreshold>'
                 '<weakClassifiers>')
        for (fi, t, l, r) in trees:
            s.append(f'<_><internalNodes>0 -1 {fi} {float(t)!r}</internalNodes>'
                     f'<leafValues>{float(l)!r} {float(r)!r}</leafValues></_>')
        s.append('</weakClassifiers></_>')
    s.append('</stages><features>')
    for rects in features:
        s.append('<_><rects>' + ''.join(f'<_>{x} {y} {ww} {hh} {float(wt)!r}</_>' for (x, y, ww, hh, wt) in rects)
                 + '</rects><tilted>0</tilted></_>')
    s.append('</features></cascade></opencv_storage>')
    with open(path, 'w') as f:
        f.write('\n'.join(s))


def random_cascade(path, rng, w=20, h=20, nstages=4, max_trees=6, nfeat=24):
    """A random multi-stage stump cascade whose stages reject roughly half of the windows each."""
    import numpy as np
    feats = []
    for _ in range(nfeat):
        nr = int(rng.integers(2, 4)); rects = []
        for k in range(nr):
            x = int(rng.integers(0, w - 2)); y = int(rng.integers(0, h - 2))
            ww = int(rng.integers(1, w - x + 1)); hh = int(rng.integers(1, h - y + 1))
            rects.append((x, y, ww, hh, float(np.float32(rng.uniform(-2, 2)))))
        feats.append(rects)
    stages = []
    for _ in range(nstages):
        nt = int(rng.integers(1, max_trees + 1))
        trees = [(int(rng.integers(0, nfeat)), float(np.float32(rng.normal(0, 0.05))),
                  float(np.float32(rng.uniform(-1, 1))), float(np.float32(rng.uniform(-1, 1)))) for _ in range(nt)]
        stages.append((float(np.float32(rng.uniform(-0.4, 0.1) * nt)), trees))
    write_cascade(path, w, h, stages, feats)


def permissive_cascade(path, rng, w, h, nstages=2, ntrees=3, nfeat=12, bias=0.35):
    """A stump cascade that accepts a large share of windows (stand-in for absent feature models)."""
    import numpy as np
    feats = []
    for _ in range(nfeat):
        rects = []
        for _k in range(2):
            x = int(rng.integers(0, w - 3)); y = int(rng.integers(0, h - 3))
            rects.append((x, y, int(rng.integers(2, w - x + 1)), int(rng.integers(2, h - y + 1)),
                          float(np.float32(rng.uniform(-2, 2)))))
        feats.append(rects)
    stages = []
    for _ in range(nstages):
        trees = [(int(rng.integers(0, nfeat)), float(np.float32(rng.normal(0, 0.05))),
                  float(np.float32(rng.uniform(-1, 1))), float(np.float32(rng.uniform(-1, 1)))) for _ in range(ntrees)]
        stages.append((float(np.float32(-bias * ntrees)), trees))
    write_cascade(path, w, h, stages, feats)


def write_old_format(path, d, name="haarcascade_converted"):
    """Writes a parsed cascade (oracle.parse_cascade_xml dict: stumps or trees, upright or tilted features) in the
    OpenCV 1.x/2.x layout: one <feature> per node; a child is <left_val>/<right_val> (leaf) or <left_node>/<right_node>."""
    s = ['<?xml version="1.0"?>', '<opencv_storage>', f'<{name} type_id="opencv-haar-classifier">',
         f'  <size>{d["win_w"]} {d["win_h"]}</size>', '  <stages>']
    t = n0 = l0 = 0
    for si, nt in enumerate(d["stage_ntrees"]):
        s.append(f'    <_>\n      <!-- stage {si} -->\n      <trees>')
        for _ in range(int(nt)):
            nn = int(d["tree_nnodes"][t])
            s.append('        <_>\n          <!-- tree -->')
            for i in range(nn):
                f = d["node_feat"][n0 + i]; R = d["feat_rect"][f]; W = d["feat_weight"][f]
                s.append('          <_>\n            <feature>\n              <rects>')
                for j in range(3):
                    if j == 2 and W[2] == 0:
                        break
                    s.append(f'                <_>{R[j][0]} {R[j][1]} {R[j][2]} {R[j][3]} {float(W[j])!r}</_>')
                s.append(f'              </rects>\n              <tilted>{int(d["feat_tilted"][f])}</tilted></feature>')
                s.append(f'            <threshold>{float(d["node_thr"][n0 + i])!r}</threshold>')
                for side, c in (("left", int(d["node_left"][n0 + i])), ("right", int(d["node_right"][n0 + i]))):
                    if c > 0:
                        s.append(f'            <{side}_node>{c}</{side}_node>')
                    else:
                        s.append(f'            <{side}_val>{float(d["leaves"][l0 - c])!r}</{side}_val>')
                s.append('          </_>')
            s.append('        </_>')
            t += 1; n0 += nn; l0 += nn + 1
        s.append(f'      </trees>\n      <stage_threshold>{float(d["stage_thr"][si])!r}</stage_threshold>\n      <parent>{si - 1}</parent>'
                 '\n      <next>-1</next></_>')
    s.append(f'  </stages></{name}>')
    s.append('</opencv_storage>')
    with open(path, 'w') as f:
        f.write('\n'.join(s))


def random_int_cascade(path, rng, w=20, h=20, nstages=12, max_trees=40, nfeat=200):
    """A random stump cascade of the kind OpenCV's trainer produces (first rect weight -1, the others small integers,
    half-rect / third-rect / checkerboard geometries): the shape the exact-integer fast kernels certify."""
    import numpy as np
    feats = []
    for _ in range(nfeat):
        kind = int(rng.integers(0, 6))
        ww = int(rng.integers(2, w + 1)); hh = int(rng.integers(2, h + 1))
        x = int(rng.integers(0, w - ww + 1)); y = int(rng.integers(0, h - hh + 1))
        if kind == 0 and ww >= 2:      # left / right half
            half = ww // 2; ww = 2 * half; side = int(rng.integers(0, 2))
            rects = [(x, y, ww, hh, -1.0), (x + side * half, y, half, hh, 2.0)]
        elif kind == 1 and hh >= 2:    # top / bottom half
            half = hh // 2; hh = 2 * half; side = int(rng.integers(0, 2))
            rects = [(x, y, ww, hh, -1.0), (x, y + side * half, ww, half, 2.0)]
        elif kind == 2 and ww >= 3:    # middle third, horizontal
            t = ww // 3; ww = 3 * t
            rects = [(x, y, ww, hh, -1.0), (x + t, y, t, hh, 3.0)]
        elif kind == 3 and hh >= 3:    # middle third, vertical
            t = hh // 3; hh = 3 * t
            rects = [(x, y, ww, hh, -1.0), (x, y + t, ww, t, 3.0)]
        elif kind == 4 and ww >= 2 and hh >= 2:   # checkerboard: three rects
            a = ww // 2; b = hh // 2; ww = 2 * a; hh = 2 * b
            rects = [(x, y, ww, hh, -1.0), (x, y, a, b, 2.0), (x + a, y + b, a, b, 2.0)]
        else:                          # an arbitrary second rect inside the first
            w2 = int(rng.integers(1, ww + 1)); h2 = int(rng.integers(1, hh + 1))
            rects = [(x, y, ww, hh, -1.0), (x + int(rng.integers(0, ww - w2 + 1)), y + int(rng.integers(0, hh - h2 + 1)), w2, h2, 2.0)]
        feats.append(rects)
    stages = []
    for s in range(nstages):
        nt = int(rng.integers(2, max_trees + 1))
        trees = [(int(rng.integers(0, nfeat)), float(np.float32(rng.normal(0, 0.02))),
                  float(np.float32(rng.uniform(-1, 1))), float(np.float32(rng.uniform(-1, 1)))) for _ in range(nt)]
        stages.append((float(np.float32(rng.uniform(-0.25, -0.05) * nt)), trees))
    write_cascade(path, w, h, stages, feats)


def random_general_model(rng, w=20, h=20, nstages=5, max_trees=8, max_nodes=3):
    """A random cascade with tree weak classifiers and (half of them) tilted features, as an oracle.parse_cascade_xml dict
    (write it with write_old_format)."""
    import numpy as np
    rects, weights, tilted, trees, stage_ntrees, stage_thr = [], [], [], [], [], []

    def feature():
        t = int(rng.integers(0, 2))
        r = np.zeros((3, 4), np.int32); wt = np.zeros(3, np.float32)
        nr = int(rng.integers(2, 4))
        for k in range(nr):
            while True:
                if t:   # tilted (x, y, w, h): x - h >= 0, x + w <= W, y + w + h <= H
                    ww = int(rng.integers(1, 8)); hh = int(rng.integers(1, 8))
                    if hh + ww > min(w, h):
                        continue
                    x = int(rng.integers(hh, w - ww + 1)); y = int(rng.integers(0, h - ww - hh + 1))
                else:
                    ww = int(rng.integers(1, w + 1)); hh = int(rng.integers(1, h + 1))
                    x = int(rng.integers(0, w - ww + 1)); y = int(rng.integers(0, h - hh + 1))
                break
            r[k] = [x, y, ww, hh]; wt[k] = np.float32([-1, 2, 3][k] if rng.integers(0, 2) else rng.uniform(-2, 2))
        rects.append(r); weights.append(wt); tilted.append(t)
        return len(rects) - 1

    for _ in range(nstages):
        nt = int(rng.integers(1, max_trees + 1))
        for _t in range(nt):
            nn = int(rng.integers(1, max_nodes + 1))
            nodes, leaves = [], []
            for i in range(nn):
                child = []
                for _side in range(2):
                    if i + 1 < nn and not any(c == i + 1 for (_, _, a, b) in nodes for c in (a, b)) and len(child) == 0 and rng.integers(0, 2):
                        child.append(i + 1)
                    else:
                        child.append(-len(leaves)); leaves.append(float(np.float32(rng.uniform(-1, 1))))
                nodes.append((feature(), float(np.float32(rng.normal(0, 0.03))), child[0], child[1]))
            # unreachable nodes are legal in the file format but make leaf counts ambiguous: keep every node reachable
            reach = {0}
            for i, (_, _, a, b) in enumerate(nodes):
                if i in reach:
                    reach.update(c for c in (a, b) if c > 0)
            if len(reach) != nn:
                nodes = nodes[:1]; leaves = [float(np.float32(rng.uniform(-1, 1))) for _ in range(2)]
                nodes[0] = (nodes[0][0], nodes[0][1], 0, -1)
            else:
                # renumber leaves in node order so that the count is nodes + 1
                k = 0; new_leaves = []; new_nodes = []
                for (f, t, a, b) in nodes:
                    ch = []
                    for c in (a, b):
                        if c > 0:
                            ch.append(c)
                        else:
                            ch.append(-k); new_leaves.append(leaves[-c]); k += 1
                    new_nodes.append((f, t, ch[0], ch[1]))
                nodes, leaves = new_nodes, new_leaves
            trees.append((nodes, leaves))
        stage_ntrees.append(nt)
        stage_thr.append(float(np.float32(rng.uniform(-0.9, -0.35) * nt)))
    import oracle as O
    return O._model(w, h, stage_ntrees, stage_thr, trees, rects, weights, tilted)


def write_lbp_cascade(path, w, h, stages, feats):
    """BOOST/LBP cascade in OpenCV's new XML layout.  stages: [(stage_thr, [tree, ...])] with tree = ([(left, right,
    feature, [8 subset words]), ...], [leaf, ...]); feats: [(x, y, cell_w, cell_h)]."""
    s = ['<?xml version="1.0"?>', '<opencv_storage>',
         '<cascade type_id="opencv-cascade-classifier"><stageType>BOOST</stageType>',
         '<featureType>LBP</featureType>', f'<height>{h}</height>', f'<width>{w}</width>',
         f'<stageParams><maxWeakCount>{max(len(t) for _, t in stages)}</maxWeakCount></stageParams>',
         '<featureParams><maxCatCount>256</maxCatCount></featureParams>', f'<stageNum>{len(stages)}</stageNum>',
         '<stages>']
    for thr, trees in stages:
        s.append(f'<_><maxWeakCount>{len(trees)}</maxWeakCount><stageThreshold>{float(thr)!r}</stageThreshold>'
                 '<weakClassifiers>')
        for nodes, leaves in trees:
            s.append('<_><internalNodes>' + ' '.join(f'{l} {r} {f} ' + ' '.join(str(int(v)) for v in sub) for (l, r, f, sub) in nodes)
                     + '</internalNodes><leafValues>' + ' '.join(repr(float(v)) for v in leaves) + '</leafValues></_>')
        s.append('</weakClassifiers></_>')
    s.append('</stages><features>')
    for (x, y, ww, hh) in feats:
        s.append(f'<_><rect>{x} {y} {ww} {hh}</rect></_>')
    s.append('</features></cascade></opencv_storage>')
    with open(path, 'w') as f:
        f.write('\n'.join(s))


def random_lbp_cascade(path, rng, w=24, h=24, nstages=6, max_trees=10, nfeat=60, max_nodes=1, pass_bias=0.0):
    """A random LBP cascade (stumps, or trees of up to max_nodes nodes) whose stages reject about half of the windows:
    random cells, random 256-bit subsets, a stage threshold near the mean of the stage sum."""
    import numpy as np
    feats = []
    for _ in range(nfeat):
        cw = int(rng.integers(1, w // 3 + 1)); ch = int(rng.integers(1, h // 3 + 1))
        feats.append((int(rng.integers(0, w - 3 * cw + 1)), int(rng.integers(0, h - 3 * ch + 1)), cw, ch))
    stages = []
    for _ in range(nstages):
        nt = int(rng.integers(1, max_trees + 1))
        trees, mean = [], 0.0
        for _t in range(nt):
            nn = int(rng.integers(1, max_nodes + 1))
            nodes, leaves = [], []
            for i in range(nn):                      # a chain: one child is the next node (if any), the other a leaf
                sub = [int(v) for v in rng.integers(-2**31, 2**31, 8)]
                if i + 1 < nn:
                    leaves.append(float(np.float32(rng.uniform(-1, 1))))
                    ch = (i + 1, -(len(leaves) - 1)) if rng.integers(0, 2) else (-(len(leaves) - 1), i + 1)
                else:
                    leaves += [float(np.float32(rng.uniform(-1, 1))) for _ in range(2)]
                    ch = (-(len(leaves) - 2), -(len(leaves) - 1))
                nodes.append((ch[0], ch[1], int(rng.integers(0, nfeat)), sub))
            trees.append((nodes, leaves))
            mean += float(np.mean(leaves))
        stages.append((float(np.float32(mean - pass_bias * nt + rng.uniform(-0.1, 0.1))), trees))
    write_lbp_cascade(path, w, h, stages, feats)
