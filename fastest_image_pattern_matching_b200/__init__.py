"""fastest_image_pattern_matching_b200 -- B200-native (sm_100a) rotation-invariant NCC template
matcher: a drop-in for the hot path of lrm2017/Fastest_Image_Pattern_Matching
(`TemplateMatcher::learnPattern` / `match`) behind the C ABI in include/fpm_b200.h."""
from .matcher import TemplateMatcher, SingleTargetMatch, FpmError, GlyphReader, match_multi, ocr_assemble  # noqa: F401
from ._build import build  # noqa: F401
