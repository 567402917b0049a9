"""Generate tests/golden/keras_model_configs.json from the reference's tests/test_model.json (the
``model.get_config()`` of ``create_model`` under TensorFlow 2.1 ... 2.5, GRU-with-attention and LSTM), which is
what the reference's own topology test compares against (tests/test_model.py:254-262).  Run in the build
container:  python tests/golden/make_model_configs.py

Kept per layer: class_name, name, inbound_nodes and the config without initializer / regularizer / constraint
entries (they do not exist at inference time)."""
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))
DROP = ("_initializer", "_regularizer", "_constraint")


def main():
    src = json.load(open("/root/reference/tests/test_model.json"))
    out = {}
    for version, kinds in src.items():
        out[version] = {}
        for kind, cfg in kinds.items():
            layers = []
            for entry in cfg["layers"]:
                conf = {k: v for k, v in entry["config"].items() if not k.endswith(DROP)}
                layers.append({"class_name": entry["class_name"], "name": entry["name"],
                               "inbound_nodes": entry["inbound_nodes"], "config": conf})
            out[version][kind] = {"name": cfg["name"], "layers": layers, "input_layers": cfg["input_layers"],
                                  "output_layers": cfg["output_layers"]}
    path = os.path.join(HERE, "keras_model_configs.json")
    with open(path, "w") as fh:
        json.dump(out, fh, indent=None, separators=(",", ":"), sort_keys=True)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
