// TEST INFRASTRUCTURE -- a plain-C doorway to the reference's own factory, compiled together with a PATCHED scratch copy
// of nbody/nbody_engines.cpp (integration/nbody_engines.patch) by tests/test_integration_patch.py:
// "engine=b200_bh;device=0;tree_layout=heap" -> nbody_create_engine(QVariantMap)
#include "nbody_engines.h"

extern "C" __attribute__((visibility("default"))) void* factory_create(const char* params)
{
	QVariantMap	m;
	QStringList	items(QString(params ? params : "").split(";", QString::SkipEmptyParts));
	for(int i = 0; i < items.size(); ++i)
	{
		int eq = items[i].indexOf("=");
		if(eq >= 0)
		{
			m[items[i].mid(0, eq).trimmed()] = QVariant(items[i].mid(eq + 1).trimmed());
		}
	}
	return nbody_create_engine(m);
}
