"""Speaker-property tokens of the controllable-TTS batch layout (data format on the caller
side of the hot path).

Same names, arguments, results and error behaviour (KeyError on an unknown category) as the reference's
`utils/properties_util.py` (`convert_properties_to_tokens` :75-81, `convert_standard_properties_to_tokens` :67-73,
`classify_speed` :83-93, `classify_pitch` :94-225).  The reference spells the pitch classes as nested if-chains; here
they are one table of upper bounds per (gender, age) searched with `bisect`, so a new statistics fit is a one-line
change.  The two places where the reference's comparisons are not what a reader expects are kept, because the token a
sample gets is part of the trained format:

* a speed of exactly 4.0 is in none of the closed-open ranges of `classify_speed` (:86-89) and lands in "very_fast";
* `GENDER_MAP` is defined twice (:24-28 and :60-63) and the second definition wins: SPCT_46 / SPCT_47, no "unknown"
  (so an unknown gender raises KeyError before the generic pitch table below could be used by
  `convert_properties_to_tokens`; `classify_pitch` itself still accepts it).
"""
from __future__ import annotations

from bisect import bisect_right

_PITCH_CLASSES = ("low_pitch", "medium_pitch", "high_pitch", "very_high_pitch")


def _numbered(prefix_first: int, names) -> dict:
    return {n: f"SPCT_{prefix_first + i}" for i, n in enumerate(names)}


SPEED_MAP = _numbered(1, ("very_slow", "slow", "medium", "fast", "very_fast"))
PITCH_MAP = _numbered(6, _PITCH_CLASSES)
AGE_MAP = _numbered(13, ("child", "teenager", "youth-adult", "middle-aged", "elderly"))
EMOTION_MAP = _numbered(21, (
    "UNKNOWN", "NEUTRAL", "ANGRY", "HAPPY", "SAD", "FEARFUL", "DISGUSTED", "SURPRISED", "SARCASTIC", "EXCITED", "SLEEPY",
    "CONFUSED", "EMPHASIS", "LAUGHING", "SINGING", "WORRIED", "WHISPER", "ANXIOUS", "NO-AGREEMENT", "APOLOGETIC",
    "CONCERNED", "ENUNCIATED", "ASSERTIVE", "ENCOURAGING", "CONTEMPT"))
GENDER_MAP = _numbered(46, ("female", "male"))

# exclusive upper bounds of the pitch classes (Hz); a table with two bounds has no "very_high_pitch" class
_PITCH_BOUNDS = {
    ("female", "child"): (250, 290),
    ("female", "teenager"): (208, 238, 270),
    ("female", "youth-adult"): (191, 211, 232),
    ("female", "middle-aged"): (176, 195, 215),
    ("female", "elderly"): (170, 190, 213),
    ("female", None): (187, 209, 232),
    ("male", "teenager"): (121, 143, 166),
    ("male", "youth-adult"): (115, 131, 153),
    ("male", "middle-aged"): (110, 125, 147),
    ("male", "elderly"): (115, 128, 142),
    ("male", None): (114, 130, 151),
    (None, None): (130, 180, 220),
}


def classify_pitch(pitch: float, gender: str, age: str) -> str:
    gender, age = gender.lower(), age.lower()
    g = gender if gender in ("female", "male") else None
    bounds = _PITCH_BOUNDS.get((g, age if g else None)) or _PITCH_BOUNDS[(g, None)]
    return _PITCH_CLASSES[bisect_right(bounds, pitch)]          # pitch < bound  <=>  bound is right of pitch


def classify_speed(speed: float) -> str:
    if speed <= 3.5:
        return "very_slow"
    if speed < 4.0:
        return "slow"
    if 4.0 < speed <= 4.5:
        return "medium"
    if 4.5 < speed <= 5.0:
        return "fast"
    return "very_fast"                                          # includes speed == 4.0, see the module docstring


def _join(age, gender, emotion, pitch_class, speed_class) -> str:
    return ("SPCT_0" + AGE_MAP[age.lower()] + GENDER_MAP[gender.lower()] + EMOTION_MAP[emotion.upper()]
            + PITCH_MAP[pitch_class] + SPEED_MAP[speed_class])


def convert_standard_properties_to_tokens(age: str, gender: str, emotion: str, pitch: str, speed: str) -> str:
    return _join(age, gender, emotion, pitch.lower(), speed.lower())


def convert_properties_to_tokens(age: str, gender: str, emotion: str, pitch: float, speed: float) -> str:
    # the reference looks the categories up before it classifies pitch and speed, so a bad category raises first
    AGE_MAP[age.lower()], GENDER_MAP[gender.lower()], EMOTION_MAP[emotion.upper()]
    return _join(age, gender, emotion, classify_pitch(pitch, gender.lower(), age.lower()), classify_speed(speed))
