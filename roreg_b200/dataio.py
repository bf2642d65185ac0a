"""Origin-data reader without open3d - SURVEY.md section 8(f) rank 3 ("ASCII-ply / keypoint loader").

The reference's dataset object (dataops/dataset.py:41-129, ThrDMatchPartDataset) is what every plugin receives: `.name`,
`.pc_ids`, `.pair_ids`, `get_kps(id)`, `get_transform(id0, id1)`.  It reads

    {root}/PointCloud/cloud_bin_{k}.ply             the scan (open3d.io.read_point_cloud, :94-95)
    {root}/PointCloud/gt.log (gtLo.log)             ground-truth pairs: 'id0 id1 n' + 4 pose rows per pair (:60-76)
    {root}/Keypoints/cloud_bin_{k}Keypoints.txt     indices of the 5000 keypoints into the scan (:113-114)

and open3d is its only dependency that is absent from this image.  `SceneFiles` is the same duck type over the same files with
a self-contained PLY reader (ASCII, binary little / big endian; any property list that contains x, y, z), so the plugins in
roreg_b200/test and roreg_b200.scene.register_scene run on the reference's data directories as they are.
Convention (dataset.py:27-30): R @ pts(id1) + t = pts(id0).
"""
import os
import numpy as np

_PLY_TYPES = {"char": "i1", "int8": "i1", "uchar": "u1", "uint8": "u1", "short": "i2", "int16": "i2", "ushort": "u2", "uint16": "u2",
              "int": "i4", "int32": "i4", "uint": "u4", "uint32": "u4", "float": "f4", "float32": "f4", "double": "f8", "float64": "f8"}


def read_ply_xyz(path):
    """Vertex positions of a PLY file as float64 [N,3] (what np.array(open3d_cloud.points) returns)."""
    with open(path, "rb") as f:
        if f.readline().strip() != b"ply":
            raise ValueError(f"{path}: not a PLY file")
        fmt, n_vertex, props, in_vertex, seen_vertex = None, None, [], False, False
        while True:
            line = f.readline()
            if not line:
                raise ValueError(f"{path}: PLY header without end_header")
            tok = line.decode("latin-1").split()
            if not tok or tok[0] in ("comment", "obj_info"):
                continue
            if tok[0] == "format":
                fmt = tok[1]
            elif tok[0] == "element":
                in_vertex = tok[1] == "vertex"
                if in_vertex:
                    if props:
                        raise ValueError(f"{path}: two vertex elements")
                    n_vertex, seen_vertex = int(tok[2]), True
                elif not seen_vertex:
                    raise ValueError(f"{path}: element '{tok[1]}' before the vertex element is not supported")
            elif tok[0] == "property" and in_vertex:
                if tok[1] == "list":
                    raise ValueError(f"{path}: list property in the vertex element")
                props.append((tok[2], _PLY_TYPES[tok[1]]))
            elif tok[0] == "end_header":
                break
        names = [p[0] for p in props]
        if n_vertex is None or not all(c in names for c in "xyz"):
            raise ValueError(f"{path}: no vertex element with x, y, z")
        if fmt == "ascii":
            cols = [names.index(c) for c in "xyz"]
            data = np.loadtxt(f, dtype=np.float64, max_rows=n_vertex, usecols=cols, ndmin=2)
            # the file stores float32 text: round through the declared type as a binary reader would
            for j, c in enumerate(cols):
                data[:, j] = data[:, j].astype(props[c][1])
            xyz = data
        elif fmt in ("binary_little_endian", "binary_big_endian"):
            end = "<" if fmt == "binary_little_endian" else ">"
            rec = np.dtype([(nm, end + ty) for nm, ty in props])
            raw = np.frombuffer(f.read(rec.itemsize * n_vertex), dtype=rec, count=n_vertex)
            xyz = np.stack([raw["x"], raw["y"], raw["z"]], 1).astype(np.float64)
        else:
            raise ValueError(f"{path}: unknown PLY format {fmt}")
    if xyz.shape != (n_vertex, 3):
        raise ValueError(f"{path}: {xyz.shape[0]} vertices read, header says {n_vertex}")
    return xyz


def parse_gt_log(path):
    """{'id0-id1': [3,4] float32} from a 3DMatch-style trajectory file: 5 lines per pair, ids on the first (any whitespace),
    then the 4 rows of the pose (the last is ignored) - dataset.py:60-76."""
    with open(path, "r") as f:
        lines = f.readlines()
    out = {}
    for k in range(len(lines) // 5):
        head = lines[5 * k].split()
        id0, id1 = int(float(head[0])), int(float(head[1]))
        rows = [np.array(lines[5 * k + 1 + r].split(), dtype=np.float32) for r in range(3)]
        out[f"{id0}-{id1}"] = np.stack(rows, 0)
    return out


class SceneFiles:
    """One scene directory of the reference's origin data (dataset.py:41-129) - same attributes and methods the plugins use."""

    def __init__(self, root_dir, stationnum, name, gt_file=None, n_keypoints=5000):
        self.root = root_dir
        self.name = name
        self.n_keypoints = n_keypoints
        self.pc_ids = [str(k) for k in range(stationnum)]
        self.pair_id2transform = parse_gt_log(gt_file or f"{root_dir}/PointCloud/gt.log")
        self.pair_ids = [tuple(v.split("-")) for v in self.pair_id2transform.keys()]
        self._kps = {}

    # the reference's accessor names (dataset.py:78-106)
    def get_pair_ids(self):
        return self.pair_ids

    def get_cloud_ids(self):
        return self.pc_ids

    def get_name(self):
        return self.name

    def get_pc_dir(self, cloud_id):
        return f"{self.root}/PointCloud/cloud_bin_{int(cloud_id)}.ply"

    def get_key_dir(self, cloud_id):
        return f"{self.root}/Keypoints/cloud_bin_{int(cloud_id)}Keypoints.txt"

    def get_pc(self, cloud_id):
        return read_ply_xyz(self.get_pc_dir(cloud_id))

    def get_transform(self, id0, id1):
        return self.pair_id2transform["-".join((id0, id1))]

    def get_kps(self, cloud_id):
        """[n_keypoints,3] float64: the scan's points at the stored keypoint indices (:111-117); without an index file the
        reference draws 5000 random points with the global NumPy RNG and stores their indices (:118-129) - same here.  Cached
        per cloud in memory (the reference re-reads the 14 MB scan on every call)."""
        cid = int(cloud_id)
        if cid in self._kps:
            return self._kps[cid]
        pc = self.get_pc(cid)
        idx_file = self.get_key_dir(cid)
        if os.path.exists(idx_file):
            idx = np.loadtxt(idx_file).astype(np.int64)
        else:
            idx = np.arange(pc.shape[0])
            np.random.shuffle(idx)
            idx = idx[0:self.n_keypoints]
            os.makedirs(os.path.dirname(idx_file), exist_ok=True)
            np.savetxt(idx_file, idx)
        self._kps[cid] = pc[idx]
        return self._kps[cid]
