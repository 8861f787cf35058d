"""Input-tuple schema shared by every model part.

careless passes its data around as one tuple of 2-D arrays whose positions are fixed by `BaseModel.input_index`
(careless/models/base.py:22-31) and read through `BaseModel.get_<column>(inputs)` static accessors (:84-121); a tuple
that reaches the `harmonic_id` slot is Laue data (:39-46).  Same names and error behaviour here, on numpy arrays;
the accessors are generated from the column list instead of being written out one by one.
"""
import numpy as np

COLUMNS = ("refl_id", "image_id", "file_id", "metadata", "intensities", "uncertainties", "wavelength", "harmonic_id")


class BaseModel:
    input_index = {name: position for position, name in enumerate(COLUMNS)}

    @staticmethod
    def get_index_by_name(name):
        try:
            return BaseModel.input_index[name]
        except KeyError:
            raise ValueError(f"name, {name}, not a valid key. Valid keys are {BaseModel.input_index.keys()}.") from None

    @staticmethod
    def get_name_by_index(index):
        if isinstance(index, (int, np.integer)) and 0 <= index < len(COLUMNS):
            return COLUMNS[index]
        raise ValueError(f"index, {index}, not a valid index. Valid indices are {BaseModel.input_index.values()}.")

    @staticmethod
    def is_laue(inputs) -> bool:
        return len(inputs) > BaseModel.input_index["harmonic_id"]

    @staticmethod
    def get_input_by_name(inputs, name):
        position = BaseModel.get_index_by_name(name)
        if position >= len(inputs):
            raise ValueError(f"Attempting to gather {name} data from input tensors with length {len(inputs)} failed.")
        column = np.asarray(inputs[position])
        # a leading batch axis of size one (a keras `fit` leftover, base.py:79-80) is dropped
        return column[0] if (column.ndim > 1 and column.shape[0] == 1) else column


def _accessor(name):
    def get(inputs):
        return BaseModel.get_input_by_name(inputs, name)
    get.__name__ = f"get_{name}"
    get.__doc__ = f"The `{name}` column of an input tuple."
    return staticmethod(get)


for _name in COLUMNS:
    setattr(BaseModel, f"get_{_name}", _accessor(_name))
