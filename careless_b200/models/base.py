"""Input-tuple schema shared by every model part (mirror of careless/models/base.py:22-121, numpy arrays)."""
import numpy as np


class BaseModel:
    input_index = {
        'refl_id': 0,
        'image_id': 1,
        'file_id': 2,
        'metadata': 3,
        'intensities': 4,
        'uncertainties': 5,
        'wavelength': 6,
        'harmonic_id': 7,
    }

    @staticmethod
    def is_laue(inputs) -> bool:
        return len(inputs) >= BaseModel.get_index_by_name('harmonic_id') + 1

    @staticmethod
    def get_name_by_index(index):
        for k, v in BaseModel.input_index.items():
            if v == index:
                return k
        raise ValueError(f"index, {index}, not a valid index. Valid indices are {BaseModel.input_index.values()}.")

    @staticmethod
    def get_index_by_name(name):
        if name not in BaseModel.input_index:
            raise ValueError(f"name, {name}, not a valid key. Valid keys are {BaseModel.input_index.keys()}.")
        return BaseModel.input_index[name]

    @staticmethod
    def get_input_by_name(inputs, name):
        idx = BaseModel.get_index_by_name(name)
        try:
            datum = inputs[idx]
        except Exception:
            raise ValueError(f"Attempting to gather {name} data from input tensors with length {len(inputs)} failed.")
        datum = np.asarray(datum)
        if datum.ndim > 1 and datum.shape[0] == 1:
            datum = np.squeeze(datum, axis=0)
        return datum

    @staticmethod
    def get_refl_id(inputs): return BaseModel.get_input_by_name(inputs, 'refl_id')
    @staticmethod
    def get_file_id(inputs): return BaseModel.get_input_by_name(inputs, 'file_id')
    @staticmethod
    def get_image_id(inputs): return BaseModel.get_input_by_name(inputs, 'image_id')
    @staticmethod
    def get_metadata(inputs): return BaseModel.get_input_by_name(inputs, 'metadata')
    @staticmethod
    def get_intensities(inputs): return BaseModel.get_input_by_name(inputs, 'intensities')
    @staticmethod
    def get_uncertainties(inputs): return BaseModel.get_input_by_name(inputs, 'uncertainties')
    @staticmethod
    def get_wavelength(inputs): return BaseModel.get_input_by_name(inputs, 'wavelength')
    @staticmethod
    def get_harmonic_id(inputs): return BaseModel.get_input_by_name(inputs, 'harmonic_id')
