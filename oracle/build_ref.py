"""Recipe that stages the REFERENCE's own Python tree as test infrastructure (oracle/_ref/, git-ignored, NOT gpurun-ignored).

The reference is pure Python, so "building" it is making its importable package and its config assets available where the
GPU box can see them (/root/reference does not exist there):

    oracle/_ref/iPERCore/...            <- /root/reference/iPERCore           (*.py only)
    oracle/_ref/assets/configs/...      <- /root/reference/assets/configs     (toml / obj-style txt / json tables)
    oracle/_ref/assets/samples/sources/donald_trump_2/00000.PNG                (the one real source image)

Nothing under oracle/_ref is tracked by git and nothing under ipercore_b200/ imports it: it is read only by tests/,
__graft_entry__.smoke() and bench.py's reference arms (oracle/ref_runtime.py).  Run by __graft_entry__.build() when
/root/reference is present; a no-op on the GPU box, which uses the staged copy that travelled with the snapshot.
"""
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
SRC = "/root/reference"


def build(force=False):
    if not os.path.isdir(os.path.join(SRC, "iPERCore")):
        return DST if os.path.isdir(os.path.join(DST, "iPERCore")) else None
    stamp = os.path.join(DST, ".staged")
    if os.path.exists(stamp) and not force:
        return DST
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    shutil.copytree(os.path.join(SRC, "iPERCore"), os.path.join(DST, "iPERCore"),
                    ignore=lambda d, names: [n for n in names if not (n.endswith(".py") or os.path.isdir(os.path.join(d, n)))])
    shutil.copytree(os.path.join(SRC, "assets", "configs"), os.path.join(DST, "assets", "configs"))
    img = os.path.join("assets", "samples", "sources", "donald_trump_2", "00000.PNG")
    if os.path.exists(os.path.join(SRC, img)):
        os.makedirs(os.path.dirname(os.path.join(DST, img)), exist_ok=True)
        shutil.copy(os.path.join(SRC, img), os.path.join(DST, img))
    open(stamp, "w").write("staged from %s\n" % SRC)
    return DST


if __name__ == "__main__":
    print(build(force=True))
