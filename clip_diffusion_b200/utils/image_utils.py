"""Range helpers of clip_diffusion/utils/image_utils.py:35-42 (the only two on the hot path).
Inside the fused cutout kernel the [-1,1] -> [0,1] map is applied on load; these torch versions
exist for callers that want the tensors themselves."""


def normalize_image_neg_one_to_one(image_tensor):
    """[0,1] -> [-1,1]  (image_utils.py:35-37)"""
    return image_tensor * 2 - 1


def denormalize_image_zero_to_one(image_tensor):
    """[-1,1] -> [0,1]  (image_utils.py:40-42)"""
    return (image_tensor + 1) / 2
