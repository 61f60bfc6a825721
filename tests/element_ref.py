"""Python restatement of the reference elements' per-frame logic on top of the CPU oracle (test infrastructure).

Each class follows one *_process_frame + *_send_event pair of the reference; line numbers as in
nubomedia-vca_b200/csrc/elements.cu.  It is written independently of the C++ mirror (list slicing instead
of iterator arithmetic, numpy float32 for the float fields) so that the two restatements check each other,
and every pixel operation goes through oracle/ (which is pinned against cv2).
"""
import math

import numpy as np

import oracle as O

F32 = np.float32


def cv_round(v):
    return int(np.rint(v))


def trunc(v):
    return int(v)


def clamp_roi(r, W, H):
    x0, y0, x1, y1 = max(r[0], 0), max(r[1], 0), min(r[0] + r[2], W), min(r[1] + r[3], H)
    if x1 <= x0 or y1 <= y0:
        return None
    return [x0, y0, x1 - x0, y1 - y0]


class Gate:
    """FACE:794-802,829-830 and its copies."""

    def __init__(self):
        self.num_frame = 0
        self.to_process = 0

    def runs(self, p):
        self.num_frame += 1
        return (p == 2 and self.num_frame % 2 == 1) or (p != 2 and self.num_frame <= p)

    def end(self):
        if self.num_frame == 4:
            self.num_frame = 0


def detect(gray, casc, sf, mn, ms):
    if casc is None or gray.shape[0] == 0 or gray.shape[1] == 0:
        return []
    return O.detect_multiscale(np.ascontiguousarray(gray), casc, sf, mn, ms).tolist()


def meta(name, typ, r):
    return (name, typ, r[0] & 0xFFFFFFFF, r[1] & 0xFFFFFFFF, r[2] & 0xFFFFFFFF, r[3] & 0xFFFFFFFF)


# ------------------------------------------------------------------------------------------------
def track_faces(faces, next_id, cur, track_threshold):
    """Faces::track_faces, FACES:78-153.  faces: [(rect, id)], cur: [rect]."""
    cf = [(list(r), i) for i, r in enumerate(cur)]
    cen = lambda r: (r[0] + r[2] // 2, r[1] + r[3] // 2)      # noqa: E731
    dist = lambda a, b: int(math.sqrt((b[0] - a[0]) ** 2 + (b[1] - a[1]) ** 2))   # noqa: E731
    out = []
    for (fr, fid) in faces:
        t, pos = track_threshold, -1
        for k, (cr, _) in enumerate(cf):
            d = dist(cen(cr), cen(fr))
            if t > d:
                pos, t = k, d
        if pos >= 0:
            cr = cf[pos][0]
            d = dist(cen(fr), cen(cr))
            a_old, a_new = fr[2] * fr[3], cr[2] * cr[3]
            big = max(a_old, a_new)
            limit = 8 if big > 5000 else (5 if big > 2500 else 3)
            if limit < d:
                out.append((cr, fid))
            elif 15 < (abs(a_old - a_new) * 100) // a_new:
                out.append(([fr[0], fr[1], cr[2], cr[3]], fid))
            else:
                out.append((fr, fid))
            del cf[pos]
    for (cr, _) in cf:
        out.append((cr, next_id))
        next_id += 1
    return out, next_id


class FaceRef:
    def __init__(self, casc):
        self.casc, self.gate = casc, Gate()
        self.faces, self.next_id, self.no_det = [], 0, 0
        self.p = dict(w2p=160, sf=25, x4=4, track=40)

    def process(self, frame):
        H, W, _ = frame.shape
        if self.gate.runs(self.p["x4"]):
            cur, _ = O.face_process(frame, self.casc, self.p["w2p"], 1.0 + self.p["sf"] / 100.0, 3, None)
            cur = cur.tolist()
            if cur:
                self.faces, self.next_id = track_faces(self.faces, self.next_id, cur, self.p["track"])
            elif self.no_det < 1:
                self.no_det += 1
            else:
                self.no_det = 0
                self.faces = []
        self.gate.end()
        norm = W // self.p["w2p"]
        return [meta("face", "face", [v * norm for v in r]) for (r, _) in self.faces]


# ------------------------------------------------------------------------------------------------
def contain_bb(px, py, r):
    return r[1] <= py <= r[1] + r[3] and r[0] <= px <= r[0] + r[2]


def merge_eyes_current_frame(face_bb, eye_r, eyes, scale, eye_left):
    """EYE:778-862 (literal, including eyes.erase(eyes.end()-i-1))."""
    area = lambda r: r[2] * r[3]      # noqa: E731
    i = len(eyes) - 1
    while i > 0:
        c = (eyes[i][0] + eyes[i][2] // 2, eyes[i][1] + eyes[i][3] // 2)
        if contain_bb(c[0], c[1], eyes[i - 1]) and area(eyes[i]) < area(eyes[i - 1]):
            del eyes[len(eyes) - i - 1]
        else:
            c = (eyes[i - 1][0] + eyes[i - 1][2] // 2, eyes[i - 1][1] + eyes[i - 1][3] // 2)
            if contain_bb(c[0], c[1], eyes[i]) and area(eyes[i - 1]) < area(eyes[i]):
                del eyes[len(eyes) - i]
        i -= 1
    i = len(eyes) - 1
    while i >= 0:
        if i < len(eyes):
            y_aux = face_bb[1] * scale + face_bb[3] * scale * 60 // 100
            if face_bb[1] * scale + eyes[i][1] < y_aux:
                if i == 0 and len(eyes) == 1:
                    if len(eye_r) > 0 and eye_left:
                        eyes[i][1] = eye_r[0][1]
                else:
                    del eyes[i]
        i -= 1
    if len(eyes) > 1:
        middle_y = face_bb[0] * scale + face_bb[3] * scale // 2
        middle_x = face_bb[1] * scale + face_bb[2] * scale // 2
        i = len(eyes) - 1
        while i > 0:
            c1 = (eyes[i][0] + eyes[i][2] // 2, eyes[i][1] + eyes[i][3] // 2)
            c2 = (eyes[i - 1][0] + eyes[i - 1][2] // 2, eyes[i - 1][1] + eyes[i - 1][3] // 2)
            s1 = F32(math.sqrt((middle_x - c1[0]) ** 2 + (middle_y - c1[1]) ** 2))
            s2 = F32(math.sqrt((middle_x - c2[0]) ** 2 + (middle_y - c2[1]) ** 2))
            if s1 < s2:
                del eyes[len(eyes) - i - 1]
            else:
                del eyes[len(eyes) - i]
            i -= 1
    if eye_left and len(eye_r) > 0 and len(eyes) > 0:
        eyes[0][1] = eye_r[0][1]


def merge_consecutive(cur, prev, limit, local, face, scale):
    """EYE:864-900 (local=False) / MOUTH:750-796, NOSE:745-790 (local=True)."""
    res = []
    for o in prev:
        oc = (o[0] + o[2] // 2, o[1] + o[3] // 2)
        for j, c in enumerate(cur):
            if local:
                nc = ((c[0] + face[0]) * scale + (c[2] * scale) // 2, (c[1] + face[1]) * scale + (c[3] * scale) // 2)
            else:
                nc = (c[0] + c[2] // 2, c[1] + c[3] // 2)
            if math.sqrt((nc[0] - oc[0]) ** 2 + (nc[1] - oc[1]) ** 2) < limit:
                res.append(list(o))
                del cur[j]
                break
    for c in cur:
        res.append([(face[0] + c[0]) * scale, (face[1] + c[1]) * scale, (c[2] - 1) * scale, (c[3] - 1) * scale] if local else list(c))
    return res


def hold(state, counter, res, max_empty):
    if not res:
        if counter < max_empty:
            return state, counter + 1
        return [], 0
    return res, 0


def scales(W, w2p, detect_event):
    o2f = F32(W) / F32(W) if detect_event else F32(W) / F32(160)
    o2x = F32(W) / F32(w2p)
    return float(o2f), float(o2x), float(F32(o2f) / F32(o2x))


class FeatureRef:
    """nuboeyedetector / nubomouthdetector / nubonosedetector on top of the oracle."""

    def __init__(self, kind, c_face, c_a, c_b=None):
        self.kind, self.c_face, self.c_a, self.c_b = kind, c_face, c_a, c_b
        self.gate = Gate()
        self.faces, self.a, self.b, self.na, self.nb = [], [], [], 0, 0
        self.p = dict(w2p=320, sf=25, x4=4, detect_event=0)
        self.queue = []

    def _receive(self):
        if not self.p["detect_event"]:
            return True
        if not self.queue:
            return False
        self.faces = [list(r) for r in self.queue.pop(0)]
        self.gate.to_process = 10 // (5 - self.p["x4"])
        return True

    def process(self, frame):
        H, W, _ = frame.shape
        o2f, o2x, f2x = scales(W, self.p["w2p"], self.p["detect_event"])
        sf = 1.0 + self.p["sf"] / 100.0
        processed = False
        if self._receive() or self.gate.to_process > 0:
            processed = True
            res_a, res_b = [], []
            ran = self.gate.runs(self.p["x4"])
            if ran:
                self.gate.to_process -= 1
                gray = O.bgr2gray(frame)
                iscale = int(o2x)
                if self.kind == "eye":
                    gray = O.equalize_hist(gray)
                    if not self.p["detect_event"]:
                        self.faces = detect(O.resize_linear(gray, cv_round(W / o2f), cv_round(H / o2f)), self.c_face, sf, 3, (30, 30))
                    feat = O.equalize_hist(O.resize_linear(gray, cv_round(W / o2x), cv_round(H / o2x)))
                    fh, fw = feat.shape
                    for f in self.faces:
                        ra = [trunc(f[0] * f2x), trunc(f[1] * f2x), trunc(f[2] * f2x), trunc(f[3] * f2x)]
                        down = cv_round(F32(ra[3]) * F32(40) / F32(100)); top = cv_round(F32(ra[3]) * F32(25) / F32(100))
                        fr = [ra[0], ra[1] + top, ra[2] // 2, ra[3] - top - down]
                        fl = [ra[0] + ra[2] // 2, ra[1] + top, ra[2] // 2, ra[3] - top - down]
                        found = []
                        for roi0, casc in ((fr, self.c_a), (fl, self.c_b)):
                            roi = clamp_roi(roi0, fw, fh)
                            ev = detect(feat[roi[1]:roi[1] + roi[3], roi[0]:roi[0] + roi[2]], casc, 1.1, 2, (20, 20)) if roi else []
                            roi = roi or roi0
                            found.append((roi, [[(roi[0] + q[0]) * iscale, (roi[1] + q[1]) * iscale, (q[2] - 1) * iscale,
                                                 (q[3] - 1) * iscale] for q in ev]))
                        (fr, eye_r), (fl, eye_l) = found
                        if eye_r:
                            merge_eyes_current_frame(fr, eye_r, eye_r, iscale, False)
                            res_a += merge_consecutive(eye_r, self.a, 7, False, fr, iscale)
                        if eye_l:
                            merge_eyes_current_frame(fl, res_a, eye_l, iscale, True)
                            res_b += merge_consecutive(eye_l, self.b, 7, False, fl, iscale)
                else:
                    if not self.p["detect_event"]:
                        small = O.equalize_hist(O.resize_linear(gray, cv_round(W / o2f), cv_round(H / o2f)))
                        self.faces = detect(small, self.c_face, sf, 2, (3, 3))
                    feat = O.equalize_hist(O.resize_linear(gray, cv_round(W / o2x), cv_round(H / o2x)))
                    fh, fw = feat.shape
                    for f in self.faces:
                        if self.kind == "mouth":
                            half = cv_round(float(F32(f[3])) / 1.8)
                            ra = [trunc(f[0] * f2x), trunc((f[1] + half) * f2x), trunc(f[2] * f2x), trunc(half * f2x)]
                        else:
                            top = cv_round(F32(f[3]) * F32(25) / F32(100)); down = cv_round(F32(f[3]) * F32(10) / F32(100))
                            side = cv_round(F32(f[2]) * F32(25) / F32(100))
                            ra = [trunc((f[0] + side) * f2x), trunc((f[1] + top) * f2x), trunc((f[2] - side) * f2x),
                                  trunc((f[3] - down - top) * f2x)]
                        roi = clamp_roi(ra, fw, fh)
                        if not roi:
                            continue
                        found = detect(feat[roi[1]:roi[1] + roi[3], roi[0]:roi[0] + roi[2]], self.c_a, 1.1, 3, (1, 1))
                        if found:
                            res_a += merge_consecutive(found, self.a, 4 if self.kind == "mouth" else 6, True, roi, iscale)
            if self.kind == "eye":
                if ran:
                    self.a, self.na = hold(self.a, self.na, res_a, 1)
                    self.b, self.nb = hold(self.b, self.nb, res_b, 1)
            else:
                self.a = res_a
            self.gate.end()
        if self.kind == "eye":
            return [meta("eye_left", "eye", r) for r in self.b] + [meta("eye_right", "eye", r) for r in self.a]
        if self.kind == "mouth":
            norm = int(o2f)
            return [meta("face", "face", [v * norm for v in f]) for f in self.faces] + [meta("mouth", "mouth", r) for r in self.a]
        return [meta("noses", "nose", r) for r in self.a]


class EarRef:
    def __init__(self, c_face, c_le, c_re):
        self.c_face, self.c_le, self.c_re = c_face, c_le, c_re
        self.gate = Gate()
        self.faces, self.lear, self.rear, self.no_det = [], [], [], 0
        self.p = dict(w2p=320, sf=25, x4=4)

    def _find(self, face_img, feat, casc, f2e, e2o, side):
        sf = 1.0 + self.p["sf"] / 100.0
        self.faces = detect(face_img, self.c_face, sf, 2, (3, 3))
        if not self.faces:
            return
        ears = self.lear if side == 0 else self.rear
        if ears:
            del ears[:]
        elif self.no_det < 4:
            self.no_det += 1
        else:
            self.no_det = 0
            del ears[:]
        fh, fw = feat.shape
        cols = face_img.shape[1]
        for f in self.faces:
            top = cv_round(F32(f[3]) * F32(20) / F32(100)); down = cv_round(F32(f[3]) * F32(20) / F32(100))
            if side == 0:
                y = trunc((f[1] + top) * f2e); x = trunc((f[0] + f[2] // 2) * f2e)
                h = trunc((f[3] - down) * f2e); w = trunc((f[2] // 2) * f2e + 50)
                if x + w > fw:
                    w = fw - x - 1
            else:
                y = trunc((f[1] + top) * f2e); x = trunc((cols - f[0] - f[2]) * f2e - 50)
                h = trunc((f[3] - down) * f2e); w = trunc((f[2] // 2) * f2e)
                if x < 0:
                    x = 0
            f[:] = [x, y, w, h]
            roi = clamp_roi(f, fw, fh)
            if not roi:
                continue
            for q in detect(feat[roi[1]:roi[1] + roi[3], roi[0]:roi[0] + roi[2]], casc, 1.1, 3, (1, 1)):
                ears.append([cv_round((roi[0] + q[0]) * e2o), cv_round((roi[1] + q[1]) * e2o), trunc((q[2] - 1) * e2o),
                             trunc((q[3] - 1) * e2o)])

    def process(self, frame):
        H, W, _ = frame.shape
        f2o = F32(W) / F32(160); e2o = F32(W) / F32(self.p["w2p"]); f2e = float(F32(f2o) / F32(e2o))
        f2o, e2o = float(f2o), float(e2o)
        if self.gate.runs(self.p["x4"]):
            gray = O.bgr2gray(frame)
            left = O.equalize_hist(O.resize_linear(gray, cv_round(W / f2o), cv_round(H / f2o)))
            feat = O.equalize_hist(O.resize_linear(gray, cv_round(W / e2o), cv_round(H / e2o)))
            self._find(left, feat, self.c_le, f2e, e2o, 0)
            self._find(np.ascontiguousarray(left[:, ::-1]), feat, self.c_re, f2e, e2o, 1)
        self.gate.end()
        msg = [meta("face_profile", "face_profile", f) for f in self.faces] + [meta("ear", "ear", r) for r in self.rear] + \
              [meta("ear", "ear", r) for r in self.lear]
        self.faces = []
        return msg
