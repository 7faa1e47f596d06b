// TEST INFRASTRUCTURE ONLY -- not part of the nb200 product.
//
// Minimal stand-in for the handful of Qt classes the drons/nbody library
// sources use (qDebug, QString, QFile/QTextStream, QVariantMap ...), so the
// reference's own engines/solvers/data sources can be compiled UNMODIFIED from
// /root/reference by oracle/Makefile in an image that has no Qt headers.
// Nothing here is derived from Qt sources: it is a from-scratch restatement
// of the documented behaviour of the few members the reference calls.
#ifndef NB200_QTSHIM_H
#define NB200_QTSHIM_H

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cctype>
#include <cmath>
#include <cassert>
#include <string>
#include <vector>
#include <map>
#include <sstream>
#include <fstream>
#include <iostream>
#include <utility>
#include <initializer_list>
#include <algorithm>
#include <functional>
#include <memory>
#include <numeric>
#include <cerrno>

// ---- qglobal.h ------------------------------------------------------------
typedef long long qint64;
typedef unsigned long long quint64;
#define Q_UNUSED(x) (void)x;
#define Q_ASSERT(c) assert(c)
#define QT_VERSION_CHECK(a, b, c) (((a) << 16) | ((b) << 8) | (c))
#define QT_VERSION QT_VERSION_CHECK(5, 12, 8)
#define Q_OBJECT
#define Q_SLOTS
#define slots

// ---- QtOpenGL (only the four names nbtype_info.h needs) ---------------------
typedef int GLint;
typedef unsigned int GLenum;
#ifndef GL_FLOAT
#define GL_FLOAT 0x1406
#define GL_DOUBLE 0x140A
#endif

namespace Qt {
enum CaseSensitivity { CaseInsensitive = 0, CaseSensitive = 1 };
}

class QChar
{
	char m_c;
public:
	QChar() : m_c(0) {}
	QChar(char c) : m_c(c) {}  // NOLINT
	char toLatin1() const { return m_c; }
};

class QStringList;

class QString
{
	std::string m_s;
public:
	enum SplitBehavior { KeepEmptyParts, SkipEmptyParts };
	QString() {}
	QString(const char* s) : m_s(s ? s : "") {}  // NOLINT
	QString(const std::string& s) : m_s(s) {}  // NOLINT
	const std::string& toStdString() const { return m_s; }
	const char* c_str() const { return m_s.c_str(); }
	bool isEmpty() const { return m_s.empty(); }
	int size() const { return static_cast<int>(m_s.size()); }
	bool operator==(const QString& o) const { return m_s == o.m_s; }
	bool operator!=(const QString& o) const { return m_s != o.m_s; }
	bool operator==(const char* o) const { return m_s == o; }
	bool operator!=(const char* o) const { return m_s != o; }
	bool operator<(const QString& o) const { return m_s < o.m_s; }
	QString operator+(const QString& o) const { return QString(m_s + o.m_s); }
	QString& operator+=(const QString& o) { m_s += o.m_s; return *this; }
	bool contains(const QString& sub) const { return m_s.find(sub.m_s) != std::string::npos; }
	int indexOf(const QString& sub) const
	{
		size_t p = m_s.find(sub.m_s);
		return p == std::string::npos ? -1 : static_cast<int>(p);
	}
	QString mid(int pos, int n = -1) const
	{
		if(pos >= size()) { return QString(); }
		return QString(n < 0 ? m_s.substr(pos) : m_s.substr(pos, n));
	}
	QString trimmed() const
	{
		size_t b = 0, e = m_s.size();
		while(b < e && isspace(static_cast<unsigned char>(m_s[b]))) { ++b; }
		while(e > b && isspace(static_cast<unsigned char>(m_s[e - 1]))) { --e; }
		return QString(m_s.substr(b, e - b));
	}
	QString repeated(int n) const
	{
		std::string r;
		for(int i = 0; i < n; ++i) { r += m_s; }
		return QString(r);
	}
	int compare(const QString& o, Qt::CaseSensitivity cs = Qt::CaseSensitive) const
	{
		if(cs == Qt::CaseSensitive) { return m_s.compare(o.m_s); }
		std::string a(m_s), b(o.m_s);
		for(auto& c : a) { c = static_cast<char>(tolower(static_cast<unsigned char>(c))); }
		for(auto& c : b) { c = static_cast<char>(tolower(static_cast<unsigned char>(c))); }
		return a.compare(b);
	}
	double toDouble(bool* ok = nullptr) const
	{
		const char* b = m_s.c_str();
		char* e = nullptr;
		double v = strtod(b, &e);
		bool good = (e != b) && (*e == 0);
		if(ok) { *ok = good; }
		return good ? v : 0.0;
	}
	unsigned long long toULongLong(bool* ok = nullptr) const
	{
		const char* b = m_s.c_str();
		char* e = nullptr;
		unsigned long long v = strtoull(b, &e, 10);
		bool good = (e != b) && (*e == 0) && m_s.find('-') == std::string::npos;
		if(ok) { *ok = good; }
		return good ? v : 0;
	}
	int toInt(bool* ok = nullptr) const
	{
		const char* b = m_s.c_str();
		char* e = nullptr;
		long v = strtol(b, &e, 10);
		bool good = (e != b) && (*e == 0);
		if(ok) { *ok = good; }
		return good ? static_cast<int>(v) : 0;
	}
	inline QStringList split(const QString& sep, SplitBehavior b = KeepEmptyParts) const;
	inline QStringList split(QChar sep, SplitBehavior b = KeepEmptyParts) const;

	// "%1" substitution, the only placeholder form the reference uses.
	QString arg(const std::string& text) const
	{
		std::string r(m_s);
		size_t p = r.find("%1");
		if(p != std::string::npos) { r.replace(p, 2, text); }
		return QString(r);
	}
	static std::string pad(std::string t, int width, QChar fill)
	{
		char f = fill.toLatin1() ? fill.toLatin1() : ' ';
		int w = width < 0 ? -width : width;
		if(static_cast<int>(t.size()) < w)
		{
			std::string padding(static_cast<size_t>(w) - t.size(), f);
			t = width < 0 ? t + padding : padding + t;
		}
		return t;
	}
	QString arg(unsigned long v, int width = 0, int base = 10, QChar fill = QChar(' ')) const
	{
		(void)base;
		return arg(pad(std::to_string(v), width, fill));
	}
	QString arg(unsigned long long v, int width = 0, int base = 10, QChar fill = QChar(' ')) const
	{
		(void)base;
		return arg(pad(std::to_string(v), width, fill));
	}
	QString arg(int v, int width = 0, int base = 10, QChar fill = QChar(' ')) const
	{
		(void)base;
		return arg(pad(std::to_string(v), width, fill));
	}
	QString arg(double v, int width = 0, char fmt = 'g', int prec = -1, QChar fill = QChar(' ')) const
	{
		char spec[16];
		char buf[128];
		snprintf(spec, sizeof(spec), "%%.%d%c", prec < 0 ? 6 : prec, fmt);
		snprintf(buf, sizeof(buf), spec, v);
		return arg(pad(buf, width, fill));
	}
	QString arg(float v, int width = 0, char fmt = 'g', int prec = -1, QChar fill = QChar(' ')) const
	{
		return arg(static_cast<double>(v), width, fmt, prec, fill);
	}
	QString arg(const QString& s) const { return arg(s.m_s); }
};

inline QString operator+(const char* a, const QString& b) { return QString(a) + b; }

class QStringList : public std::vector<QString>
{
public:
	int size() const { return static_cast<int>(std::vector<QString>::size()); }
	bool isEmpty() const { return empty(); }
};

inline QStringList QString::split(const QString& sep, SplitBehavior b) const
{
	QStringList out;
	size_t pos = 0;
	const std::string& d = sep.m_s;
	while(true)
	{
		size_t n = d.empty() ? std::string::npos : m_s.find(d, pos);
		std::string part = m_s.substr(pos, n == std::string::npos ? std::string::npos : n - pos);
		if(!(b == SkipEmptyParts && part.empty())) { out.push_back(QString(part)); }
		if(n == std::string::npos) { break; }
		pos = n + d.size();
	}
	return out;
}

inline QStringList QString::split(QChar sep, SplitBehavior b) const
{
	return split(QString(std::string(1, sep.toLatin1())), b);
}

template<class T>
class QVector : public std::vector<T>
{
public:
	QVector() {}
	void resize(int n) { std::vector<T>::resize(static_cast<size_t>(n)); }
	int size() const { return static_cast<int>(std::vector<T>::size()); }
};

template<class A, class B>
struct QPair
{
	A first;
	B second;
	QPair() : first(), second() {}
	QPair(const A& a, const B& b) : first(a), second(b) {}
};
template<class A, class B>
QPair<A, B> qMakePair(const A& a, const B& b) { return QPair<A, B>(a, b); }

// ---- QVariant / QVariantMap -----------------------------------------------
class QVariant
{
	bool		m_valid;
	std::string	m_s;
public:
	QVariant() : m_valid(false) {}
	QVariant(const char* s) : m_valid(true), m_s(s) {}  // NOLINT
	QVariant(const QString& s) : m_valid(true), m_s(s.toStdString()) {}  // NOLINT
	QVariant(const std::string& s) : m_valid(true), m_s(s) {}  // NOLINT
	QVariant(bool v) : m_valid(true), m_s(v ? "true" : "false") {}  // NOLINT
	QVariant(int v) : m_valid(true), m_s(std::to_string(v)) {}  // NOLINT
	QVariant(unsigned v) : m_valid(true), m_s(std::to_string(v)) {}  // NOLINT
	QVariant(long v) : m_valid(true), m_s(std::to_string(v)) {}  // NOLINT
	QVariant(unsigned long v) : m_valid(true), m_s(std::to_string(v)) {}  // NOLINT
	QVariant(unsigned long long v) : m_valid(true), m_s(std::to_string(v)) {}  // NOLINT
	QVariant(double v) : m_valid(true)  // NOLINT
	{
		char buf[64];
		snprintf(buf, sizeof(buf), "%.17g", v);
		m_s = buf;
	}
	bool isValid() const { return m_valid; }
	QString toString() const { return QString(m_s); }
	double toDouble() const { return QString(m_s).toDouble(); }
	int toInt() const { return static_cast<int>(toDouble()); }
	unsigned toUInt() const { return static_cast<unsigned>(toDouble()); }
	unsigned long long toULongLong() const { return static_cast<unsigned long long>(toDouble()); }
	bool toBool() const
	{
		QString l(m_s);
		return m_valid && !m_s.empty() && l.compare("false", Qt::CaseInsensitive) != 0 && m_s != "0";
	}
};

template<class K, class V>
class QMap : public std::map<K, V>
{
public:
	QMap() {}
	QMap(const std::map<K, V>& m) : std::map<K, V>(m) {}  // NOLINT
	V value(const K& k, const V& def = V()) const
	{
		auto it = this->find(k);
		return it == this->end() ? def : it->second;
	}
	bool contains(const K& k) const { return this->find(k) != this->end(); }
};
typedef QMap<QString, QVariant> QVariantMap;

// ---- QDebug ----------------------------------------------------------------
// Writes one space-separated line to stderr when the last copy dies; silenced
// by NBREF_QUIET=1 so test logs stay readable.
class QDebug
{
	struct state
	{
		std::ostringstream	os;
		int					refs;
	};
	state* m_st;
	void sep() { m_st->os << ' '; }
public:
	QDebug() : m_st(new state) { m_st->refs = 1; }
	QDebug(const QDebug& o) : m_st(o.m_st) { ++m_st->refs; }
	QDebug& operator=(const QDebug&) = delete;
	~QDebug()
	{
		if(--m_st->refs == 0)
		{
			static const bool quiet = (getenv("NBREF_QUIET") != nullptr);
			if(!quiet) { fprintf(stderr, "%s\n", m_st->os.str().c_str()); }
			delete m_st;
		}
	}
	QDebug& noquote() { return *this; }
	QDebug& nospace() { return *this; }
	QDebug& operator<<(const char* s) { m_st->os << (s ? s : "(null)"); sep(); return *this; }
	QDebug& operator<<(const QString& s) { m_st->os << s.toStdString(); sep(); return *this; }
	QDebug& operator<<(const std::string& s) { m_st->os << s; sep(); return *this; }
	QDebug& operator<<(const QVariant& s) { m_st->os << s.toString().toStdString(); sep(); return *this; }
	QDebug& operator<<(bool v) { m_st->os << (v ? "true" : "false"); sep(); return *this; }
	QDebug& operator<<(char v) { m_st->os << v; sep(); return *this; }
	QDebug& operator<<(int v) { m_st->os << v; sep(); return *this; }
	QDebug& operator<<(unsigned v) { m_st->os << v; sep(); return *this; }
	QDebug& operator<<(long v) { m_st->os << v; sep(); return *this; }
	QDebug& operator<<(unsigned long v) { m_st->os << v; sep(); return *this; }
	QDebug& operator<<(long long v) { m_st->os << v; sep(); return *this; }
	QDebug& operator<<(unsigned long long v) { m_st->os << v; sep(); return *this; }
	QDebug& operator<<(float v) { m_st->os << v; sep(); return *this; }
	QDebug& operator<<(double v) { m_st->os << v; sep(); return *this; }
	QDebug& operator<<(long double v) { m_st->os << v; sep(); return *this; }
	QDebug& operator<<(const void* p) { m_st->os << p; sep(); return *this; }
	template<class T>
	QDebug& operator<<(const std::vector<T>& v)
	{
		m_st->os << '(';
		for(size_t i = 0; i < v.size(); ++i) { if(i) { m_st->os << ", "; } m_st->os << v[i]; }
		m_st->os << ')';
		sep();
		return *this;
	}
	QDebug& operator<<(const QStringList& v)
	{
		m_st->os << '(';
		for(int i = 0; i < v.size(); ++i) { if(i) { m_st->os << ", "; } m_st->os << v[i].toStdString(); }
		m_st->os << ')';
		sep();
		return *this;
	}
	QDebug& operator<<(const QVariantMap& m)
	{
		m_st->os << "QMap(";
		for(auto& kv : m) { m_st->os << '(' << kv.first.toStdString() << ", " << kv.second.toString().toStdString() << ')'; }
		m_st->os << ')';
		sep();
		return *this;
	}
};
inline QDebug qDebug() { return QDebug(); }

// ---- QFile / QTextStream ---------------------------------------------------
class QFile
{
	std::string		m_name;
	std::fstream	m_f;
	std::string		m_err;
public:
	enum OpenMode { ReadOnly = 1, WriteOnly = 2 };
	explicit QFile(const QString& name) : m_name(name.toStdString()) {}
	bool open(OpenMode mode)
	{
		m_f.open(m_name.c_str(), mode == ReadOnly ? std::ios::in : (std::ios::out | std::ios::trunc));
		if(!m_f.is_open()) { m_err = strerror(errno); return false; }
		return true;
	}
	QString errorString() const { return QString(m_err); }
	std::fstream& stream() { return m_f; }
};

class QTextStream
{
	QFile*	m_file;
	int		m_precision;
	bool	m_scientific;
	bool	m_force_sign;
public:
	enum RealNumberNotation { SmartNotation, FixedNotation, ScientificNotation };
	enum NumberFlag { ShowBase = 1, ForcePoint = 2, ForceSign = 4 };
	explicit QTextStream(QFile* f) : m_file(f), m_precision(6), m_scientific(false), m_force_sign(false) {}
	void setRealNumberPrecision(int p) { m_precision = p; }
	void setRealNumberNotation(RealNumberNotation n) { m_scientific = (n == ScientificNotation); }
	void setNumberFlags(int flags) { m_force_sign = (flags & ForceSign) != 0; }
	bool atEnd() { return m_file->stream().peek() == std::char_traits<char>::eof(); }
	QString readLine()
	{
		std::string line;
		std::getline(m_file->stream(), line);
		if(!line.empty() && line.back() == '\r') { line.pop_back(); }
		return QString(line);
	}
	QTextStream& operator<<(double v)
	{
		char spec[16];
		char buf[128];
		snprintf(spec, sizeof(spec), "%%%s.%d%c", m_force_sign ? "+" : "", m_precision, m_scientific ? 'e' : 'g');
		snprintf(buf, sizeof(buf), spec, v);
		m_file->stream() << buf;
		return *this;
	}
	QTextStream& operator<<(float v) { return *this << static_cast<double>(v); }
	QTextStream& operator<<(const char* s) { m_file->stream() << s; return *this; }
	QTextStream& operator<<(const QString& s) { m_file->stream() << s.toStdString(); return *this; }
};

#endif // NB200_QTSHIM_H
