"""Evaluation metrics of the reference's losses.py:276-366 (MOSI ``full_loss``, IEMOCAP
``iemocap_loss``, POM ``pom_loss``).  Host-side NumPy / scikit-learn reporting, outside the
hot path; kept so that ``from losses import full_loss, ...`` works and the
downstream parity check ("MAE and correlation unchanged to the 3rd decimal") uses the same
definitions: same keys, same rounding, same label conventions.
"""
import numpy as np
from sklearn.metrics import accuracy_score, classification_report, confusion_matrix, f1_score


def _binary_report(true_label, predicted_label):
    conf = confusion_matrix(true_label, predicted_label)
    print("Confusion Matrix :")
    print(conf)
    print("Classification Report :")
    print(classification_report(true_label, predicted_label, digits=5))
    return conf, classification_report(true_label, predicted_label, digits=5, output_dict=True)


def full_loss(predictions, y_test):
    """reference losses.py:276-316 -- MOSI: MAE, Pearson r, 7-class accuracy / weighted F1 on
    rounded scores (5 dp), binary accuracy on the sign (>= 0)."""
    p = np.asarray(predictions).flatten()
    y = np.asarray(y_test).flatten()
    mae = np.mean(np.absolute(p - y))
    corr = np.corrcoef(p, y)[0][1]
    mult = round(sum(np.round(p) == np.round(y)) / float(len(y)), 5)
    f_score = round(f1_score(np.round(p), np.round(y), average='weighted'), 5)
    print("mae: {}".format(mae))
    print("corr: {}".format(corr))
    print("mult_acc: {}".format(mult))
    print("mult f_score: {}".format(f_score))
    true_label, predicted_label = (y >= 0), (p >= 0)
    accuracy = accuracy_score(true_label, predicted_label)
    conf, report = _binary_report(true_label, predicted_label)
    print("Accuracy {}".format(accuracy))
    return {'mae': float(mae), 'accuracy': float(accuracy), 'corr': float(corr), 'mult_acc': float(mult),
            'f_score': float(f_score), 'confusion_matrix': conf.tolist(), 'class_report': report}


def iemocap_loss(predictions, y_test):
    """reference losses.py:318-342 -- arg-max class accuracy and weighted F1."""
    true_label = np.argmax(y_test, axis=1)
    predicted_label = np.argmax(predictions, axis=1)
    f_score = f1_score(true_label, predicted_label, average='weighted')
    accuracy = accuracy_score(true_label, predicted_label)
    print("F1 score:", f_score)
    print("Accuracy:", accuracy)
    conf, report = _binary_report(true_label, predicted_label)
    return {'accuracy': float(accuracy), 'f_score': float(f_score), 'confusion_matrix': conf.tolist(),
            'class_report': report}


def pom_loss(predictions, y_test):
    """reference losses.py:344-366 -- per-trait MAE / r / rounded accuracy (3 dp), F1 (5 dp)."""
    predictions, y_test = np.asarray(predictions), np.asarray(y_test)
    n_traits = y_test.shape[1]
    mae = [round(a, 3) for a in np.mean(np.absolute(predictions - y_test), axis=0)]
    corr = [round(np.corrcoef(predictions[:, i], y_test[:, i])[0][1], 3) for i in range(n_traits)]
    mult = [round(sum(np.round(predictions[:, i]) == np.round(y_test[:, i])) / float(len(y_test)), 3)
            for i in range(n_traits)]
    f_score = [round(f1_score(np.round(predictions[:, i]), np.round(y_test[:, i]), average='weighted'), 5)
               for i in range(n_traits)]
    print("mae:", mae)
    print("corr:", corr)
    print('mult_acc:', mult)
    print('f_score:', f_score)
    return {'mae': [float(x) for x in mae], 'corr': [float(x) for x in corr],
            'mult_acc': [float(x) for x in mult], 'f_score': [float(x) for x in f_score]}
