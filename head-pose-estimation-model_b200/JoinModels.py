"""``join_models`` with the contract of the reference's JoinModels.py:5-90: attach two regressor
heads to the BlazeFace detector at ``re_lu_10`` / ``re_lu_15`` and save the unified model.

Here the join is a spec merge (no Keras graph surgery): detector weights + two head specs ->
``UnifiedModel`` whose six outputs keep the reference order (JoinModels.py:152-158)."""
import os

from . import h5lite
from .keras_spec import load_model
from .unified import UnifiedModel, _normalise, pack_backbone

TAP_CHANNELS = {"re_lu_10": 88, "re_lu_15": 96}


def join_models(face_detector_path, regressor1_path, regressor2_path, layer1_name="re_lu_10", layer2_name="re_lu_15",
                output_model_path=None, metadata: dict = None):
    for path in (face_detector_path, regressor1_path, regressor2_path):
        if not os.path.exists(path):
            raise FileNotFoundError(f"Model file not found: {path}")
    for lname in (layer1_name, layer2_name):
        if lname not in TAP_CHANNELS:
            raise ValueError(f"Layer '{lname}' not found in face detector model")
    if (layer1_name, layer2_name) != ("re_lu_10", "re_lu_15"):
        raise ValueError("regressor1 attaches to 're_lu_10' and regressor2 to 're_lu_15' (JoinModels.py:117-118)")
    det = _normalise(h5lite.H5File(face_detector_path).weights())
    pack_backbone(det)  # validates that this is a BlazeFace-front detector (raises ValueError otherwise)
    reg1, reg2 = load_model(regressor1_path), load_model(regressor2_path)
    unified = UnifiedModel(det, reg1, reg2)
    unified._metadata = metadata or {}
    if output_model_path:
        unified.save(output_model_path)
    return unified


def extract_id_from_path(file_path):
    return os.path.basename(file_path)[:-3] if file_path.endswith(".h5") else None
