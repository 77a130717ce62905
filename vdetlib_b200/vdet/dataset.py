"""ImageNet VID class table (reference vdet/dataset.py:7-11 reads misc/imagenet_vdet_classes.txt).

Index 0 is the background; the "30 VID classes" of BASELINE.json are indices 1..30.
"""
imagenet_vdet_classes = [
    "__background__", "airplane", "antelope", "bear", "bicycle", "bird", "bus", "car", "cattle",
    "dog", "domestic_cat", "elephant", "fox", "giant_panda", "hamster", "horse", "lion", "lizard",
    "monkey", "motorcycle", "rabbit", "red_panda", "sheep", "snake", "squirrel", "tiger", "train",
    "turtle", "watercraft", "whale", "zebra",
]
imagenet_vdet_class_idx = dict(zip(imagenet_vdet_classes, range(len(imagenet_vdet_classes))))
