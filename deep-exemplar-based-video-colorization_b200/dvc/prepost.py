"""Host-side geometry of the reference's image ingest (test.py:44-46): CenterPad (utils/util_distortion.py:217-258) followed
by torchvision CenterCrop, reduced to ONE device call -- dvc_resize_antialias_crop_rgb8(src, (Hr, Wr), (oy, ox), (Ho, Wo)):
out[y, x] = resize(src, (Hr, Wr))[y + oy, x + ox] where that pixel exists, 0 elsewhere.

The integer arithmetic below is the reference's own (Python float division + int() truncation), so e.g. a 1080x1920 frame
with image_size (432, 768) gives exactly the size the reference computes."""


def _center_crop_offsets(h, w, th, tw):
    """torchvision.transforms.functional.center_crop: (pad_top, pad_left, crop_top, crop_left)."""
    pt = (th - h) // 2 if th > h else 0
    pl = (tw - w) // 2 if tw > w else 0
    pb = (th - h + 1) // 2 if th > h else 0
    pr = (tw - w + 1) // 2 if tw > w else 0
    h2, w2 = h + pt + pb, w + pl + pr
    if (h2, w2) == (th, tw) and (pt or pl or pb or pr):
        return pt, pl, 0, 0
    return pt, pl, int(round((h2 - th) / 2.0)), int(round((w2 - tw) / 2.0))


def centerpad_geometry(height_old, width_old, size):
    """-> (Hr, Wr, oy, ox) for an input of height_old x width_old and target size = (height, width)."""
    height, width = size
    old_size = [height_old, width_old]
    ratio = height / width
    if height_old / width_old == ratio:
        if height_old == height:
            Hr, Wr = height_old, width_old
        else:
            Hr, Wr = [int(x * height / height_old) for x in old_size]
        oy = ox = 0
        ph, pw = Hr, Wr
    elif height_old / width_old > ratio:  # pad the width and crop (util_distortion.py:244-250)
        Hr, Wr = [int(x * width / width_old) for x in old_size]
        oy, ox = (Hr - height) // 2, 0
        ph, pw = height, width
        if Wr != width:
            raise ValueError("CenterPad: the reference itself fails here (resized width != target width)")
    else:  # pad the height and crop (util_distortion.py:251-257)
        Hr, Wr = [int(x * height / height_old) for x in old_size]
        oy, ox = 0, (Wr - width) // 2
        ph, pw = height, width
        if Hr != height:
            raise ValueError("CenterPad: the reference itself fails here (resized height != target height)")
    pt, pl, ct, cl = _center_crop_offsets(ph, pw, height, width)
    return Hr, Wr, oy + ct - pt, ox + cl - pl
